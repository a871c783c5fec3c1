// Hand-written sm_100a kernels of the openpose-plus post-processing path.
//
//   K2  k2_peaks_fast<S,R,STORE>   Gaussian smoothing + 3x3 max NMS + threshold      (src/post-process.h:51-111,155-203)
//                                  + raster-order peak list                           (src/post-process.h:190-198,205-213)
//                                  + (STORE) the x8 / x4 up-sampling of all 57 maps   (src/post-process.h:24-49) written
//                                    from inside the same kernel; exact skipping of blocks provably below the threshold
//       k2_peaks_generic           the same stage for any kernel size / any scale, on a materialised heat map
//   K1  k1_replicate_chw / _hwc    stand-alone up-sampling, channels-first (streaming stores) and channels-last (TMA
//       k1_general                 bulk stores); table-driven 2x2-tap version for non-integer scales
//   K3  k3_limbs                   PAF line-integral scoring of candidate pairs      (src/paf.cpp:79-134,313-337)
//                                  std::sort order + greedy bipartite matching       (src/paf.cpp:136-175)
//                                  person assembly + human_t output (last limb CTA)  (src/paf.cpp:177-262,292-310)
//   K0  k0_ingest, k0_hwc_to_chw   input staging helpers
//
// All arithmetic that decides an integer output is written with explicit round-to-nearest
// intrinsics in the reference's operation order (the library is also built with -fmad=false), so
// results are bit-identical to a strict-IEEE build of the reference.  No tensor cores: nothing here
// is a dense contraction.  Citations are relative to /root/reference.
//
// Debug / experiment switches (environment, read once): OPP_TRACE (event + phase-stamp timelines),
// OPP_K2_NOSKIP, OPP_NO_FUSE, OPP_FORCE_GENERIC, OPP_K2_TW / OPP_K2_TH (tile size), OPP_K1_MODE
// (0 persistent lean stores, 1 TMA bulk stores, 2 = default streaming stores), OPP_K1_G, OPP_K1_CTAS,
// OPP_NO_ZEROCOPY_OUT, OPP_INGEST_MAX, OPP_ZC_IN_MAX, OPP_NO_PRIORITY.
#include "opp_kernels.cuh"

#include <math_constants.h>

#include <cstdint>
#include <cstdlib>

namespace
{
// include/openpose-plus/coco.h:11-53
__constant__ int c_pair_a[OPP_N_PAIRS] = {1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5};
__constant__ int c_pair_b[OPP_N_PAIRS] = {2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17};
__constant__ int c_net_x[OPP_N_PAIRS] = {12, 20, 14, 16, 22, 24, 0, 2, 4, 6, 8, 10, 28, 30, 34, 32, 36, 18, 26};
// first limb whose (a, b) contains the part: that limb's CTA writes the part's slice of all_peaks
__constant__ int c_part_writer[OPP_N_PARTS] = {12, 0, 0, 2, 3, 1, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 14, 16};

__device__ __forceinline__ int reflect101(int p, int n)
{
    // cv::BORDER_REFLECT_101; a single reflection suffices because radius < n is enforced on the host
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

// Programmatic dependent launch (latency path, opp_capi.cu): a kernel launched with the programmatic-serialization
// attribute may be scheduled while its predecessor still runs; it must not touch the predecessor's results before
// pdl_wait() (which returns once the predecessor grid has completed and flushed).  pdl_trigger() in the predecessor
// lets that early scheduling begin.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ int clip_idx(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

// Bounds checking for the arrays the kernels carve out of shared memory (and a few global ones).  compute-sanitizer is
// not available on the GPU pool this is developed on, so the library can be built a second time with -DOPP_DEBUG_BOUNDS
// (openpose_plus_b200/build.py build_debug -> libopp_b200_dbg.so): every Span access is then checked, the first
// violation is recorded (source line of the Span's declaration, index, size) and the access is redirected to element 0
// instead of corrupting memory; opp_debug_bounds_report() reads the record.  In the release build a Span is a bare pointer.
#ifdef OPP_DEBUG_BOUNDS
__device__ int g_oob[4]; // line, index, size, number of violations
__device__ __noinline__ void oob_record(int line, long i, long n)
{
    if (atomicAdd(&g_oob[3], 1) == 0) g_oob[0] = line, g_oob[1] = (int)i, g_oob[2] = (int)n;
}
template <typename T> struct Span {
    T *p;
    long n;
    int line;
    __device__ __forceinline__ T &operator[](long i) const
    {
        if (i < 0 || i >= n) oob_record(line, i, n), i = 0;
        return p[i];
    }
    __device__ __forceinline__ Span<T> from(long ofs) const { return Span<T>{p + ofs, n - ofs, line}; } // the tail starting at ofs
};
#define SPAN(T, ptr, n) Span<T> { (ptr), (long)(n), __LINE__ }
#else
template <typename T> struct Span {
    T *p;
    __device__ __forceinline__ T &operator[](long i) const { return p[i]; }
    __device__ __forceinline__ Span<T> from(long ofs) const { return Span<T>{p + ofs}; }
};
#define SPAN(T, ptr, n) Span<T> { (ptr) }
#endif

// Asynchronous global -> shared staging (cp.async / LDGSTS): every thread queues all its pieces
// before anything waits, so a tile costs one memory round trip instead of one per loop iteration.
// 16-byte pieces when both sides are 16-byte aligned, 4-byte pieces otherwise.  Completed by
// stage_wait() followed by __syncthreads().
__device__ __forceinline__ void stage_async(float *sdst, const float *gsrc, int n)
{
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sdst);
    if (((reinterpret_cast<uintptr_t>(gsrc) | sbase) & 15) == 0) {
        const int n4 = n >> 2;
        for (int t = threadIdx.x; t < n4; t += blockDim.x)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * t), "l"(gsrc + 4 * t) : "memory");
        for (int t = 4 * n4 + threadIdx.x; t < n; t += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + 4u * t), "l"(gsrc + t) : "memory");
    } else {
        for (int t = threadIdx.x; t < n; t += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + 4u * t), "l"(gsrc + t) : "memory");
    }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// TMA bulk load (cp.async.bulk global -> shared, completion on an mbarrier): ONE instruction moves a
// whole contiguous tile; the copy engine does the rest while the CTA sets up.  Requires 16-byte
// aligned source, destination and size.  bulk_load() is called by a single thread after
// bulk_init(); every thread that reads the tile calls bulk_wait() first.
__device__ __forceinline__ void bulk_init(unsigned long long *mbar)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(float *sdst, const float *gsrc, unsigned bytes, unsigned long long *mbar)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(mbar), d = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(a)
                 : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned long long *mbar, unsigned phase)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(phase) : "memory");
    } while (!ok);
}

// One sample of the area-mode up-sampled map (cv::resize INTER_AREA, dst >= src): horizontal 2-tap
// on the two source rows, then vertical 2-tap.  Same float operations as the row-buffer form.
__device__ __forceinline__ float upsample_at(const OppGeom &g, const float *plane, int y, int x)
{
    if (g.S > 0) return plane[(y / g.S) * g.w + (x / g.S)];
    const int sy = g.yofs[y];
    const float *S0 = plane + clip_idx(sy, g.h) * g.w;
    const float *S1 = plane + clip_idx(sy + 1, g.h) * g.w;
    const int sx = g.xofs[x];
    float r0, r1;
    if (x < g.xmax) {
        const float a0 = g.alpha[2 * x], a1 = g.alpha[2 * x + 1];
        r0 = __fadd_rn(__fmul_rn(S0[sx], a0), __fmul_rn(S0[sx + 1], a1));
        r1 = __fadd_rn(__fmul_rn(S1[sx], a0), __fmul_rn(S1[sx + 1], a1));
    } else {
        r0 = __fmul_rn(S0[sx], 1.f);
        r1 = __fmul_rn(S1[sx], 1.f);
    }
    return __fadd_rn(__fmul_rn(r0, g.beta[2 * y]), __fmul_rn(r1, g.beta[2 * y + 1]));
}

// ------------------------------------------------------------------------------------------------
// K1: resize.  Store-bound: 4*C*H*W bytes written per frame against 4*C*h*w read.
// ------------------------------------------------------------------------------------------------

// Integer scale, CHW, W % 4 == 0, S % 4 == 0: every float4 is one source value replicated.
// One CTA per (group of source rows, channel, frame); each warp streams whole output rows with
// 16-byte stores, a warp instruction covering 512 contiguous bytes.
template <int S>
__global__ void __launch_bounds__(OPP_THREADS) k1_replicate_chw(const K1Params p, int rows_per_cta)
{
    extern __shared__ float s_src[]; // [rows_per_cta][w]
    const int w = p.g.w, W = p.g.W, h = p.g.h;
    const int f = blockIdx.z;
    int c = blockIdx.y, C = p.C;
    const float *src_base = p.src;
    float *dst_base = p.dst;
    if (c >= p.C) c -= p.C, C = p.C2, src_base = p.src2, dst_base = p.dst2;
    const int i0 = blockIdx.x * rows_per_cta;
    const int i1 = min(i0 + rows_per_cta, h);
    const float *src = src_base + ((size_t)f * C + c) * h * w;
    for (int t = threadIdx.x; t < (i1 - i0) * w; t += blockDim.x) s_src[t] = __ldg(src + i0 * w + t);
    __syncthreads();
    float4 *dst = reinterpret_cast<float4 *>(dst_base + ((size_t)f * C + c) * p.g.H * W);
    const int W4 = W >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int y = i0 * S + warp; y < i1 * S; y += nwarps) {
        const float *row = s_src + (y / S - i0) * w;
        float4 *drow = dst + (size_t)y * W4;
        for (int v = lane; v < W4; v += 32) {
            const float s = row[(v * 4) / S];
            __stcs(drow + v, make_float4(s, s, s, s));
        }
    }
}

// Lean per-thread stores: one CTA per (channel, frame) plane; thread = (float4 column v, half of the S
// replicated rows).  Per feature row a thread loads one value and issues S/2 16-byte streaming stores
// through pre-computed row pointers: ~3.5 instructions per 512-byte warp store.
template <int S>
__global__ void __launch_bounds__(224) k1_replicate_rows(const K1Params p, int G, int n_items)
{
    // Persistent: a fixed, small number of CTAs per SM (chosen by the launcher) walks over the
    // (row group, channel, frame) items, so the store stream keeps a bounded footprint and the
    // compute-bound kernels it overlaps with always find room on every SM.
    const int w = p.g.w, W = p.g.W, h = p.g.h, H = p.g.H;
    const int W4 = W >> 2;
    const int groups = (h + G - 1) / G;
    const int Ctot = p.C + p.C2;
    constexpr int HALF = S / 2;
    const int t = threadIdx.x;
    if (t >= 2 * W4) return;
    const int half = t / W4, v = t - half * W4;
    const int sofs = (v * 4) / S;
    const size_t dofs = (size_t)(half * HALF) * W4 + v;
    const size_t step = (size_t)S * W4;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int gi = item % groups;
        const int t2 = item / groups;
        int c = t2 % Ctot, C = p.C;
        const int f = t2 / Ctot;
        const float *src_base = p.src;
        float *dst_base = p.dst;
        if (c >= p.C) c -= p.C, C = p.C2, src_base = p.src2, dst_base = p.dst2;
        const int i0 = gi * G, i1 = min(i0 + G, h);
        const float *sp = src_base + ((size_t)f * C + c) * h * w + (size_t)i0 * w + sofs;
        float4 *dp = reinterpret_cast<float4 *>(dst_base + ((size_t)f * C + c) * H * W) + (size_t)i0 * S * W4 + dofs;
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
            const float sv = __ldg(sp);
            sp += w;
            const float4 val = make_float4(sv, sv, sv, sv);
#pragma unroll
            for (int j = 0; j < HALF; ++j) __stcs(dp + (size_t)j * W4, val);
            dp += step;
        }
    }
}

// TMA bulk-store variant (the default): a CTA expands G feature rows into shared memory once
// (16-byte shared stores), then ONE thread hands each expanded row to the copy engine S times
// (cp.async.bulk shared -> global, one per replicated output row).  The SM issues ~7 instructions per
// KB written instead of ~90 with per-thread stores, so the store stream leaves the issue slots to the
// compute-bound peak / limb kernels it runs beside.  Two shared buffers: the rows of item k+1 are
// expanded while the copy engine still drains item k.
__device__ __forceinline__ void bulk_store_row(float *gdst, const float *ssrc, unsigned bytes)
{
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}

template <int S>
__global__ void __launch_bounds__(128) k1_replicate_bulk(const K1Params p, int G, int n_items)
{
    extern __shared__ __align__(128) float sbuf[]; // [2][G][W]
    const int w = p.g.w, W = p.g.W, h = p.g.h, H = p.g.H;
    const int groups = (h + G - 1) / G;
    const int Ctot = p.C + p.C2;
    const int W4 = W >> 2;
    int buf = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
        const int gi = item % groups;
        const int t2 = item / groups;
        int c = t2 % Ctot, C = p.C;
        const int f = t2 / Ctot;
        const float *src_base = p.src;
        float *dst_base = p.dst;
        if (c >= p.C) c -= p.C, C = p.C2, src_base = p.src2, dst_base = p.dst2;
        const int i0 = gi * G, i1 = min(i0 + G, h);
        const float *src = src_base + ((size_t)f * C + c) * h * w + (size_t)i0 * w;
        float *dst = dst_base + ((size_t)f * C + c) * H * W;
        float *sb = sbuf + (size_t)buf * G * W;
        // the copies issued from this buffer two items ago must have finished READING it (bulk groups
        // are per thread: every issuing thread waits on its own)
        const bool issuer = threadIdx.x < G * S;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int t = threadIdx.x; t < (i1 - i0) * W4; t += blockDim.x) {
            const int r = t / W4, v = t - r * W4;
            const float sv = __ldg(src + r * w + (v * 4) / S);
            reinterpret_cast<float4 *>(sb)[r * W4 + v] = make_float4(sv, sv, sv, sv);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the copy engine
        __syncthreads();
        if (issuer) {
            const int r = threadIdx.x / S, q = threadIdx.x - r * S;
            if (r < i1 - i0) bulk_store_row(dst + ((size_t)(i0 + r) * S + q) * W, sb + (size_t)r * W, (unsigned)(W * sizeof(float)));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x < G * S) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the reads
}

// Channels-last output ([n, H, W, C], what the Python PostProcessor returns) at an integer scale: one
// CTA per (feature row, tensor, frame) builds ONE output row (W pixels x C channels, 33 KB for the heat
// maps and 66 KB for the PAFs at 432 columns) in shared memory - the S image rows of a feature row are
// identical - and hands it to the copy engine S times (TMA bulk store, one cp.async.bulk per row).
template <int S>
__global__ void __launch_bounds__(OPP_THREADS) k1_replicate_hwc(const K1Params p)
{
    extern __shared__ __align__(128) float srow[]; // [W*C] output row, then [C][w] feature row
    const int w = p.g.w, W = p.g.W, h = p.g.h, H = p.g.H;
    const int i = blockIdx.x, f = blockIdx.z;
    const bool second = blockIdx.y == 1;
    const int C = second ? p.C2 : p.C;
    const float *src = (second ? p.src2 : p.src) + (size_t)f * C * h * w + (size_t)i * w;
    float *dst = (second ? p.dst2 : p.dst) + ((size_t)f * H + (size_t)i * S) * W * C;
    float *feat = srow + (size_t)W * C;
    for (int t = threadIdx.x; t < C * w; t += blockDim.x) {
        const int c = t / w, x = t - c * w;
        feat[t] = __ldg(src + (size_t)c * h * w + x);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < W * C; t += blockDim.x) {
        const int x = t / C, c = t - x * C;
        srow[t] = feat[c * w + x / S];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < S) {
        const unsigned bytes = (unsigned)((size_t)W * C * sizeof(float));
        bulk_store_row(dst + (size_t)threadIdx.x * W * C, srow, bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must outlive the reads
    }
}

// Any geometry / layout: one thread per output element, table-driven 2x2-tap sample.
__global__ void __launch_bounds__(OPP_THREADS) k1_general(const K1Params p)
{
    const size_t per_frame = (size_t)p.C * p.g.H * p.g.W;
    const size_t total = per_frame * p.n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / per_frame);
        size_t r = i - (size_t)f * per_frame;
        int c, y, x;
        if (p.layout == OPP_LAYOUT_CHW) {
            x = (int)(r % p.g.W), r /= p.g.W;
            y = (int)(r % p.g.H), c = (int)(r / p.g.H);
        } else {
            c = (int)(r % p.C), r /= p.C;
            x = (int)(r % p.g.W), y = (int)(r / p.g.W);
        }
        const float *plane = p.src + ((size_t)f * p.C + c) * p.g.h * p.g.w;
        __stcs(p.dst + i, upsample_at(p.g, plane, y, x));
    }
}

// Any scale, channels-first: cv::resize's own structure.  One CTA per (block of output rows, plane, frame): the feature
// rows under the block are resized horizontally ONCE into shared memory (cv's row buffers: S[sx] a0 + S[sx+1] a1), every
// output row is then the vertical blend of two of them (r0 b0 + r1 b1), written with coalesced streaming stores.
// Same operations on the same operands as upsample_at; store-bound like the replication kernels.
#define K1G_ROWS 32
__global__ void __launch_bounds__(OPP_THREADS) k1_general_chw(const K1Params p)
{
    extern __shared__ __align__(16) float smem[];
    const OppGeom &g = p.g;
    const int W = g.W, H = g.H, h = g.h, w = g.w;
    int c = blockIdx.y;
    const float *src = p.src;
    float *dst = p.dst;
    int C = p.C;
    if (c >= p.C) c -= p.C, src = p.src2, dst = p.dst2, C = p.C2;
    const int f = blockIdx.z, y_a = blockIdx.x * K1G_ROWS, y_b = min(y_a + K1G_ROWS, H);
    const float *plane = src + ((size_t)f * C + c) * h * w;
    float *out = dst + ((size_t)f * C + c) * (size_t)H * W;
    const int f_lo = clip_idx(g.yofs[y_a], h), f_hi = clip_idx(g.yofs[y_b - 1] + 1, h), nfr = f_hi - f_lo + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // row buffers: a warp per feature row, lanes across it (no per-element division)
    for (int fr = warp; fr < nfr; fr += nwarps) {
        const float *S0 = plane + (f_lo + fr) * w;
        float *Trow = smem + fr * W;
        for (int x = lane; x < W; x += 32) {
            const int sx = g.xofs[x];
            Trow[x] = x < g.xmax ? __fadd_rn(__fmul_rn(S0[sx], g.alpha[2 * x]), __fmul_rn(S0[sx + 1], g.alpha[2 * x + 1])) : __fmul_rn(S0[sx], 1.f);
        }
    }
    __syncthreads();
    // vertical blend: a warp per output row (source rows and coefficients once per row), 16-byte streaming stores
    const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int y = y_a + warp; y < y_b; y += nwarps) {
        const int sy = g.yofs[y];
        const float b0 = g.beta[2 * y], b1 = g.beta[2 * y + 1];
        const float *R0 = smem + (clip_idx(sy, h) - f_lo) * W, *R1 = smem + (clip_idx(sy + 1, h) - f_lo) * W;
        float *orow = out + (size_t)y * W;
        if (vec) {
            for (int x4 = lane; x4 < (W >> 2); x4 += 32) {
                const float4 r0 = reinterpret_cast<const float4 *>(R0)[x4], r1 = reinterpret_cast<const float4 *>(R1)[x4];
                float4 v;
                v.x = __fadd_rn(__fmul_rn(r0.x, b0), __fmul_rn(r1.x, b1));
                v.y = __fadd_rn(__fmul_rn(r0.y, b0), __fmul_rn(r1.y, b1));
                v.z = __fadd_rn(__fmul_rn(r0.z, b0), __fmul_rn(r1.z, b1));
                v.w = __fadd_rn(__fmul_rn(r0.w, b0), __fmul_rn(r1.w, b1));
                __stcs(reinterpret_cast<float4 *>(orow) + x4, v);
            }
        } else {
            for (int x = lane; x < W; x += 32) __stcs(orow + x, __fadd_rn(__fmul_rn(R0[x], b0), __fmul_rn(R1[x], b1)));
        }
    }
}

// Latency-mode ingest: a few frames in pinned host memory are pulled over PCIe by the SMs (wide
// coalesced loads, every request in flight at once) into the staging buffers, and the frame
// counters are cleared by the same launch: one kernel instead of a memset and two DMA copies whose
// fixed costs dominate at this size.
__global__ void __launch_bounds__(OPP_THREADS) k0_ingest(const float *__restrict__ src0, float *__restrict__ dst0, size_t n0,
                                                         const float *__restrict__ src1, float *__restrict__ dst1, size_t n1, int *counters, int n_counters)
{
    pdl_trigger();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < (size_t)n_counters; i += nth) counters[i] = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(src1) | reinterpret_cast<uintptr_t>(dst0) | reinterpret_cast<uintptr_t>(dst1)) & 15) == 0 &&
                     ((n0 | n1) & 3) == 0;
    if (vec) {
        const size_t m0 = n0 >> 2, m1 = n1 >> 2;
        for (size_t i = tid; i < m0 + m1; i += nth) {
            if (i < m0)
                reinterpret_cast<float4 *>(dst0)[i] = __ldcs(reinterpret_cast<const float4 *>(src0) + i);
            else
                reinterpret_cast<float4 *>(dst1)[i - m0] = __ldcs(reinterpret_cast<const float4 *>(src1) + (i - m0));
        }
    } else {
        for (size_t i = tid; i < n0 + n1; i += nth) {
            if (i < n0)
                dst0[i] = __ldcs(src0 + i);
            else
                dst1[i - n0] = __ldcs(src1 + (i - n0));
        }
    }
}

__global__ void __launch_bounds__(OPP_THREADS) k0_hwc_to_chw(const float *src, float *dst, int n, int C, int hw)
{
    const size_t total = (size_t)n * C * hw;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int px = (int)(i % hw);
        const size_t r = i / hw;
        const int c = (int)(r % C);
        const size_t f = r / C;
        dst[i] = __ldg(src + (f * hw + px) * C + c);
    }
}

// ------------------------------------------------------------------------------------------------
// The peak kernels leave one unordered list of keys (y * W + x) per (frame, part).  The reference's all_peaks vector
// is those peaks in raster order part -> y -> x with the running index as id (src/post-process.h:176-199): a peak's
// rank inside its part's list plus the sizes of the parts before it.  The limb kernel's CTAs establish that order,
// each for the two parts of its limb (every part belongs to at least one limb): order_part_peaks ranks one part's
// keys (staged in shared memory), hands the (x, y) of every peak to the caller in raster order and - in the first limb
// that contains the part (c_part_writer) - writes the part's slice of the global all_peaks array.
// ------------------------------------------------------------------------------------------------
struct PeakSource {
    OppGeom g;
    const float *conf;    // [n,19,h,w] feature maps: score = up-sampled, unsmoothed value at the peak (src/post-process.h:192)
    const float *conf_up; // materialised [n,19,H,W] map when the geometry has no closed form (non-integer scale)
};

__device__ __forceinline__ void order_part_peaks(const PeakSource &src, int frame, int part, const Span<int> s_keys, int n, int ofs, const Span<int2> s_xy,
                                                 opp_peak_t *out /* frame's all_peaks, or nullptr: another limb writes this part's slice */)
{
    const int W = src.g.W, H = src.g.H;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int key = s_keys[t];
        int rank = 0;
        for (int u = 0; u < n; ++u) rank += (s_keys[u] < key);
        const int y = key / W, x = key - y * W;
        s_xy[rank] = make_int2(x, y);
        if (!out) continue;
        float score;
        if (src.conf_up)
            score = __ldcg(src.conf_up + ((size_t)frame * OPP_N_HEAT + part) * H * W + key);
        else
            score = upsample_at(src.g, src.conf + ((size_t)frame * OPP_N_HEAT + part) * src.g.h * src.g.w, y, x);
        opp_peak_t pk;
        pk.part_id = part, pk.x = x, pk.y = y, pk.score = score, pk.id = ofs + rank;
        out[ofs + rank] = pk;
    }
}

__device__ __forceinline__ bool tile_done_is_last(int *counter, int total)
{
    // Every CTA publishes its writes with a release-increment (ordered after the whole CTA's writes by
    // the barrier); only the CTA that sees the final count pays for the acquire side (L1 invalidate).
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        int old;
        asm volatile("atom.add.release.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        s_last = (old == total - 1);
        if (s_last) __threadfence();
    }
    __syncthreads();
    return s_last != 0;
}

__device__ __forceinline__ void emit_peak(const K2Params &p, int frame, int part, int y, int x)
{
    const int slot = atomicAdd(p.cnt.pk_cnt + frame * OPP_N_PARTS + part, 1);
    if (slot < p.capP) p.pk_key[((size_t)frame * OPP_N_PARTS + part) * p.capP + slot] = y * p.g.W + x;
}

// ------------------------------------------------------------------------------------------------
// K2 fast path: integer scale S, Gaussian radius R <= 2S.
//
// The up-sampled map is S x S blocks of one feature value, so
//   * the row pass of cv::GaussianBlur only has h distinct input rows (not H): it is evaluated once
//     per feature row into shared memory (Rrow[h][W]);
//   * inside one feature row/column the 2R+1 taps touch only three distinct neighbours (a, b, c) when
//     R <= S and five (a2, a, b, c, c2) when S < R <= 2S (k = 19..33 at x8: the Python graph's fixed
//     k = 25 among them), so the tap products are shared between the S phases.  Every product and
//     every sum is still the same IEEE operation on the same operands in the same order as OpenCV's
//     scalar filter (row: s = k0*x0; s += kj*xj left to right;  column: s = kR*x; s += k(R+j)*(x[+j] + x[-j])),
//     so the smoothed map is bit-identical; it never exists in HBM.
// The border rule only changes WHICH neighbour a tap reads (make_win): BORDER_REFLECT_101 at the image
// edge maps onto the same neighbours except for the single taps at distance exactly S and 2S (the *s
// operands); the Python graph's 'SAME' zero padding (BZ) makes the cells beyond the edge read 0.
// tests/test_fast_kernel_model.py is the executable CPU model of exactly this operand scheme.
// ------------------------------------------------------------------------------------------------
struct Win9 {
    float a2s, a2, as, a, b, c, cs, c2, c2s;
};

// Operands for cell i of n along one axis: (m2, m1, z, p1, p2) = cells i-2 .. i+2 (anything where the cell does
// not exist: such a value is never selected).  NB = neighbour depth (1: R <= S, 2: S < R <= 2S; needs n >= NB + 1).
template <int NB, bool BZ>
__device__ __forceinline__ Win9 make_win(float m2, float m1, float z, float p1, float p2, int i, int n)
{
    Win9 v;
    const bool lo1 = i > 0, hi1 = i < n - 1;
    v.b = z;
    if (BZ) {
        v.a = v.as = lo1 ? m1 : 0.f;
        v.c = v.cs = hi1 ? p1 : 0.f;
    } else {
        v.a = lo1 ? m1 : z, v.as = lo1 ? m1 : p1;
        v.c = hi1 ? p1 : z, v.cs = hi1 ? p1 : m1;
    }
    if (NB > 1) {
        const bool lo2 = i > 1, hi2 = i < n - 2;
        if (BZ) {
            v.a2 = v.a2s = lo2 ? m2 : 0.f;
            v.c2 = v.c2s = hi2 ? p2 : 0.f;
        } else {
            v.a2 = lo2 ? m2 : (lo1 ? m1 : p1), v.a2s = lo2 ? m2 : (lo1 ? z : p2);
            v.c2 = hi2 ? p2 : (hi1 ? p1 : m1), v.c2s = hi2 ? p2 : (hi1 ? z : m2);
        }
    } else {
        v.a2 = v.a2s = v.c2 = v.c2s = 0.f;
    }
    return v;
}

// value of the replicated line at offset o from the start of the cell (o in [-2S, 3S-1]; o is a
// compile-time constant at every call site once the loops are unrolled)
template <int S>
__device__ __forceinline__ float win_src(const Win9 &v, int o)
{
    if (o >= 0 && o < S) return v.b;
    if (o < 0) {
        const int d = -o;
        return d < S ? v.a : (d == S ? v.as : (d < 2 * S ? v.a2 : v.a2s));
    }
    const int d = o - S + 1;
    return d < S ? v.c : (d == S ? v.cs : (d < 2 * S ? v.c2 : v.c2s));
}

template <int S, int R, bool BZ>
__device__ __forceinline__ void row_taps(const float *__restrict__ k, const Win9 &v, float (&out)[S])
{
#pragma unroll
    for (int q = 0; q < S; ++q) {
        float s;
        if (R == 1 && !BZ) { // OpenCV's SymmRowSmallFilter, ksize 3
            s = __fadd_rn(__fmul_rn(win_src<S>(v, q), k[1]), __fmul_rn(__fadd_rn(win_src<S>(v, q - 1), win_src<S>(v, q + 1)), k[2]));
        } else if (R == 2 && !BZ) { // ksize 5
            s = __fadd_rn(__fmul_rn(win_src<S>(v, q), k[2]), __fmul_rn(__fadd_rn(win_src<S>(v, q - 1), win_src<S>(v, q + 1)), k[3]));
            s = __fadd_rn(s, __fmul_rn(__fadd_rn(win_src<S>(v, q - 2), win_src<S>(v, q + 2)), k[4]));
        } else { // RowFilter: taps left to right
            s = 0.f;
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) {
                const float pr = __fmul_rn(k[j], win_src<S>(v, q + j - R));
                s = (j == 0) ? pr : __fadd_rn(s, pr);
            }
        }
        out[q] = s;
    }
}

template <int S, int R>
__device__ __forceinline__ float col_phase(const float *__restrict__ k, const int ph, const Win9 &v)
{
    float s = __fmul_rn(k[R], v.b);
#pragma unroll
    for (int j = 1; j <= R; ++j) s = __fadd_rn(s, __fmul_rn(k[R + j], __fadd_rn(win_src<S>(v, ph + j), win_src<S>(v, ph - j))));
    return s;
}

template <int S, int R>
__device__ __forceinline__ void col_all(const float *__restrict__ k, const Win9 &v, float (&s)[S])
{
#pragma unroll
    for (int ph = 0; ph < S; ++ph) s[ph] = col_phase<S, R>(k, ph, v);
}

// The column pass of a thread's TWO image columns at once with sm_100's packed FP32 adds (add.rn.f32x2, SASS FADD2:
// two independent IEEE additions per instruction).  The additions are two thirds of the filter's arithmetic and the
// peak kernel is FP32-issue bound whenever most blocks are active.  Every lane of a packed add is the same
// round-to-nearest addition on the same operands as the scalar form; the products stay scalar multiplications, so
// there is no multiply feeding a packed add inside one instruction stream that ptxas could contract into FFMA2
// (build.py greps the SASS for FFMA2 / FFMA to make sure).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// RowFilter (taps left to right) for two neighbouring outputs per packed add: the same scalar products, the same
// additions in the same order, two chains per instruction.
template <int S, int R>
__device__ __forceinline__ void row_taps2(const float *__restrict__ k, const Win9 &v, float (&out)[S])
{
    static_assert(S % 2 == 0, "output pairs");
#pragma unroll
    for (int q = 0; q < S; q += 2) {
        f32x2_t s = pack2(__fmul_rn(k[0], win_src<S>(v, q - R)), __fmul_rn(k[0], win_src<S>(v, q + 1 - R)));
#pragma unroll
        for (int j = 1; j <= 2 * R; ++j) s = add2(s, pack2(__fmul_rn(k[j], win_src<S>(v, q + j - R)), __fmul_rn(k[j], win_src<S>(v, q + 1 + j - R))));
        unpack2(s, out[q], out[q + 1]);
    }
}

template <int S, int R>
__device__ __forceinline__ void col_all2(const float *__restrict__ k, const Win9 &v0, const Win9 &v1, float (&s0)[S], float (&s1)[S])
{
#pragma unroll
    for (int ph = 0; ph < S; ++ph) {
        f32x2_t s = pack2(__fmul_rn(k[R], v0.b), __fmul_rn(k[R], v1.b));
#pragma unroll
        for (int j = 1; j <= R; ++j) {
            const f32x2_t ps = add2(pack2(win_src<S>(v0, ph + j), win_src<S>(v1, ph + j)), pack2(win_src<S>(v0, ph - j), win_src<S>(v1, ph - j)));
            float p0, p1;
            unpack2(ps, p0, p1);
            s = add2(s, pack2(__fmul_rn(k[R + j], p0), __fmul_rn(k[R + j], p1)));
        }
        unpack2(s, s0[ph], s1[ph]);
    }
}

#ifndef K2_FAST_MAX_THREADS
#define K2_FAST_MAX_WARPS K2_FAST_MAX_GROUPS // column groups per CTA: 7 x 62 decided columns >= 432
#define K2_FAST_MAX_THREADS (32 * K2_FAST_MAX_GROUPS)
#define K2_FAST_MIN_CTAS 3
#endif

// State of the running 3x3 max for one image column: horizontal 3-max of the two previous rows and
// the smoothed value of the previous row (the one being decided).
struct NmsState {
    float pp_h, p_h, p_s;
};

// STORE = true additionally materialises the up-sampled maps from inside this kernel (the resize of
// src/post-process.h:24-49 fused with its consumer): grid.y grows to 19 and the CTA of "part" k
// streams heat map k and PAF channels 2k, 2k+1 of its tile to HBM, 12 16-byte stores per thread and
// feature row, interleaved with the filter arithmetic of the same rows.  The store stream then shares
// SMs with the FP32 work by construction instead of relying on two kernels being co-scheduled (a
// stand-alone resize next to this kernel serialises: both want every SM).  CTA 18 (background heat
// map, PAF 36/37) only stores.
__device__ __forceinline__ void stamp2(const K2Params &p, int slot)
{
    if (p.times && threadIdx.x == 0 && blockIdx.z == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.times[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + slot] = t;
    }
}

template <int S, int R, bool STORE, bool BZ>
__global__ void __launch_bounds__(K2_FAST_MAX_THREADS, STORE ? K2_FAST_MIN_CTAS : K2_FAST_MIN_CTAS + 1) k2_peaks_fast(const K2Params p)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int NB = R > S ? 2 : 1; // neighbour depth: cells a tap can reach on either side
    static_assert(R <= 2 * S, "Gaussian radius beyond two feature cells");
    const int frame = blockIdx.z, part = blockIdx.y;
    stamp2(p, 0);
    const bool compute = part < OPP_N_PARTS;
    const int tx = blockIdx.x % p.nxs, ty = blockIdx.x / p.nxs;
    const int h = p.g.h, w = p.g.w, H = S * h;
    const int ja = tx * p.tw, jb = min(ja + p.tw, w);
    const int ia = ty * p.th, ib = min(ia + p.th, h);
    const int jlo = max(ja - NB, 0), jhi = min(jb + NB, w);         // cells of the activity masks and of Rrow
    const int ilo = max(ia - 1 - NB, 0), ihi = min(ib + 1 + NB, h); // halo row + its own filter neighbours
    const int nr = ihi - ilo, ncol = jhi - jlo;
    const int RW = S * ncol;
    const Span<float> L = SPAN(float, smem, nr * w);                                       // [nr][w] feature rows
    const Span<float> Rrow = SPAN(float, smem + ((nr * w + 3) & ~3), nr * RW);             // [nr][RW] row-pass result
    const Span<float> Pf = SPAN(float, Rrow.p + ((nr * RW + 3) & ~3), 2 * (ib - ia) * w);  // [2][ib-ia][w] PAF feature rows of the tile (STORE only)
    const float *src = p.conf + ((size_t)frame * OPP_N_HEAT + part) * h * w + ilo * w;
    pdl_trigger(); // the limb kernel may be scheduled behind this grid at once (it can fetch PAF tiles from pinned memory
                   // meanwhile); it waits for our completion itself before touching the peaks
    pdl_wait();    // the maps (and the counters) may come from an ingest kernel still in flight
    stage_async(L.p, src, nr * w);
    if (STORE) {
        // staged up front: a load issued next to the store stream would queue behind it for microseconds
        const int nt = (ib - ia) * w;
        const float *ps = p.paf + (((size_t)frame * OPP_N_PAF + 2 * part) * h + ia) * w;
        stage_async(Pf.p, ps, nt);
        stage_async(Pf.p + nt, ps + (size_t)h * w, nt);
    }
    stage_wait();
    __syncthreads();
    stamp2(p, 1);

    // ---- store plan (STORE): thread = (16-byte column v of the tile, half of the S replicated rows).
    // Units = (plane, feature row) in plane-major order (heat tile, then the two PAF tiles), so each
    // CTA writes ONE contiguous stream at a time (DRAM pages); per unit a thread writes HALF rows of
    // its column.  Units are issued all along the CTA's life - a few between the prologue phases,
    // three per feature row of the column loop - so the store stream never pauses for a whole phase.
    constexpr int HALF = S / 2;
    const int X0 = S * ja, X1 = S * jb;
    const int W4 = (S * w) >> 2, Wt4 = (X1 - X0) >> 2;
    const bool st_on = STORE && (int)threadIdx.x < 2 * Wt4;
    const int nrt = ib - ia, st_total = 3 * nrt;
    float4 *ob_h = nullptr, *ob_p = nullptr;
    int fcol = 0;
    const size_t plane4 = (size_t)H * W4;
    if (st_on) {
        const int half = (int)threadIdx.x / Wt4, v = (int)threadIdx.x - half * Wt4;
        fcol = ja + (4 * v) / S;
        const size_t tofs = (size_t)(S * ia + half * HALF) * W4 + (X0 >> 2) + v;
        ob_h = reinterpret_cast<float4 *>(p.up_conf) + ((size_t)frame * OPP_N_HEAT + part) * plane4 + tofs;
        ob_p = reinterpret_cast<float4 *>(p.up_paf) + ((size_t)frame * OPP_N_PAF + 2 * part) * plane4 + tofs;
    }
    int st_unit = 0;
    auto store_units = [&](int count) {
        if (st_on) {
            for (int rep = 0; rep < count && st_unit < st_total; ++rep, ++st_unit) {
                const int plane = st_unit / nrt, r = st_unit - plane * nrt;
                const float sv = plane == 0 ? L[(ia - ilo + r) * w + fcol] : Pf[((plane - 1) * nrt + r) * w + fcol];
                float4 *o = (plane == 0 ? ob_h : ob_p + (size_t)(plane - 1) * plane4) + (size_t)r * S * W4;
                const float4 val = make_float4(sv, sv, sv, sv);
#pragma unroll
                for (int j = 0; j < HALF; ++j) __stcs(o + (size_t)j * W4, val);
            }
        }
    };
    const int st_pro = STORE ? (st_total + 7) / 8 : 0; // units issued in each of the three prologue gaps
    store_units(st_pro);


    // ---- which blocks can hold a peak at all?  A pixel of feature cell (r, c) is filtered from the
    // (2 NB + 1)^2 cells around it (R <= NB * S, reflection included; cells beyond a zero border only lower it).  Every float operation of the filter is monotone
    // in its inputs and the taps are positive, so if those nine values are all <= t then the smoothed
    // pixel is <= t * (sum of taps, rounded up) < t * (1 + 2^-17).  With t = thresh * (1 - 2^-13) the
    // pixel cannot exceed thresh, cannot be a peak, and cannot outrank a pixel that is one: its block
    // (column group x feature row) need not be computed and its pixels may stand as -inf for their
    // neighbours.  Exact for every input; dense maps simply have every block active.
    __shared__ unsigned long long s_act[64], s_need[64], s_blk[K2_FAST_MAX_WARPS];
    const int ngroups = (X1 - X0 + 61) / 62;
    if (compute) {
        const float t_skip = p.skip_thresh;
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, nwarps_ = blockDim.x >> 5;
        // (1) cells above the skip threshold, one ballot per 32 cells (a warp per feature row) ...
        for (int r = warp_; r < nr; r += nwarps_) {
            unsigned long long bits = 0ull;
            for (int c0 = 0; c0 < ncol; c0 += 32) {
                const bool hot = c0 + lane_ < ncol && L[r * w + jlo + c0 + lane_] > t_skip;
                bits |= (unsigned long long)__ballot_sync(0xffffffffu, hot) << c0;
            }
            // only decided cells [ja, jb) are ever tested, and their neighbourhoods lie inside the
            // cells [jlo, jhi) and the staged rows (or end at the image border)
            if (lane_ == 0) s_act[r] = bits;
        }
        __syncthreads();
        // (2) active blocks: warp g owns column group g; lane = feature row; the dilation of the bit rows
        //     is done on the fly (rows r-NB .. r+NB OR-ed, then shifted left and right by up to NB cells)
        if (warp_ < ngroups) {
            const int g = warp_;
            const int xa = X0 + 62 * g, xb = min(xa + 62, X1) - 1;
            const int ca = xa / S - jlo, cb = xb / S - jlo;
            const unsigned long long dec = (cb - ca >= 63 ? ~0ull : ((1ull << (cb - ca + 1)) - 1)) << ca; // cells of the decided columns
            unsigned long long blk = 0ull;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int r = lane_ + 32 * k;
                unsigned long long v = 0ull;
                if (r < nr) {
                    v = s_act[r];
                    if (r > 0) v |= s_act[r - 1];
                    if (r + 1 < nr) v |= s_act[r + 1];
                    if (NB > 1) {
                        if (r > 1) v |= s_act[r - 2];
                        if (r + 2 < nr) v |= s_act[r + 2];
                    }
                    v |= (v << 1) | (v >> 1) | (NB > 1 ? (v << 2) | (v >> 2) : 0ull);
                }
                blk |= (unsigned long long)__ballot_sync(0xffffffffu, (v & dec) != 0) << (32 * k);
            }
            if (lane_ == 0) s_blk[g] = blk;
        }
        __syncthreads();
        // (3) the cells the row pass must produce: every cell a live block reads (its 64 columns, rows r-NB..r+NB)
        for (int r = threadIdx.x; r < nr; r += blockDim.x) {
            unsigned long long need = 0ull;
            const unsigned long long span = (1ull << (2 * NB + 1)) - 1;
            const unsigned long long near3 = r >= NB ? span << (r - NB) : span >> (NB - r); // rows r-NB .. r+NB
            for (int g = 0; g < ngroups; ++g) {
                if (s_blk[g] & near3) {
                    const int xa = X0 + 62 * g, xb = min(xa + 62, X1) - 1;
                    const int sa = max(xa - 1, S * jlo) / S - jlo, sb = min(xb + 1, S * jhi - 1) / S - jlo;
                    need |= (sb - sa >= 63 ? ~0ull : ((1ull << (sb - sa + 1)) - 1)) << sa;
                }
            }
            s_need[r] = need;
        }
        __syncthreads();
    }

    stamp2(p, 2);
    store_units(st_pro);
    // ---- row pass: one thread per (feature row, feature column) -> S outputs
    if (compute) {
        for (int it = threadIdx.x; it < nr * ncol; it += blockDim.x) {
            const int r = it / ncol, c = jlo + (it - r * ncol);
            if (!((s_need[r] >> (c - jlo)) & 1ull)) continue; // no active block reads this cell
            const Span<float> Lr = L.from(r * w);
            float out[S];
            const float m2 = NB > 1 ? Lr[max(c - 2, 0)] : 0.f, m1 = Lr[max(c - 1, 0)], p1 = Lr[min(c + 1, w - 1)], p2 = NB > 1 ? Lr[min(c + 2, w - 1)] : 0.f;
            constexpr bool kRowPacked = R > 2 || BZ; // the symmetric-small forms (k = 3, 5) stay as OpenCV writes them
            const Win9 win = make_win<NB, BZ>(m2, m1, Lr[c], p1, p2, c, w);
            if (kRowPacked) row_taps2<S, R>(p.taps, win, out);
            else row_taps<S, R, BZ>(p.taps, win, out);
            float *d = &Rrow[r * RW + S * (c - jlo)];
            if (S % 4 == 0) {
#pragma unroll
                for (int q = 0; q < S; q += 4) *reinterpret_cast<float4 *>(d + q) = make_float4(out[q], out[q + 1], out[q + 2], out[q + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < S; ++q) d[q] = out[q];
            }
        }
    }
    __syncthreads();
    stamp2(p, 3);
    store_units(st_pro);

    const int xlo = S * jlo, xhi = S * jhi; // columns present in Rrow
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float NINF = -CUDART_INF_F;

    if (!compute || warp >= ngroups) { // store-only CTA (background) or a warp without a column group
        if (STORE) store_units(st_total);
        if (!compute) return;
    } else {
        // ---- column pass + 3x3 max + threshold.  A warp walks down 64 image columns, two per lane: lane l
        // holds columns xg0-1+2l and xg0+2l, so 62 columns are decided per warp and the outermost two only
        // feed their neighbours.  Columns that do not exist read -inf, which the filter maps to -inf.
        const int x0 = X0 + 62 * warp - 1 + 2 * lane, x1 = x0 + 1;
        const bool have0 = x0 >= xlo && x0 < xhi, have1 = x1 >= xlo && x1 < xhi;
        const float thr0 = (lane >= 1 && x0 < X1) ? p.thresh : CUDART_INF_F;
        const float thr1 = (lane <= 30 && x1 < X1) ? p.thresh : CUDART_INF_F;
        const Span<float> col0 = Rrow.from(have0 ? x0 - xlo : 0), col1 = Rrow.from(have1 ? x1 - xlo : 0);
        NmsState n0 = {NINF, NINF, NINF}, n1 = {NINF, NINF, NINF};

        // one finished image row: decide the row above it.  Decisions are collected as bit masks (bit =
        // phase of the finished row) and emitted once per feature row: peaks are rare, and one
        // branch per 8 rows keeps the arithmetic of the 16 running sums in one basic block.
        unsigned m0 = 0, m1 = 0;
        auto step = [&](float s0, float s1, int bit, bool decide) {
            const float l = __shfl_up_sync(0xffffffffu, s1, 1);
            const float r = __shfl_down_sync(0xffffffffu, s0, 1);
            const float h0 = fmaxf(fmaxf(l, s0), s1), h1 = fmaxf(fmaxf(s0, s1), r);
            if (decide) {
                const bool e0 = n0.p_s > thr0 && n0.p_s >= fmaxf(fmaxf(n0.pp_h, n0.p_h), h0);
                const bool e1 = n1.p_s > thr1 && n1.p_s >= fmaxf(fmaxf(n1.pp_h, n1.p_h), h1);
                m0 |= e0 ? (1u << bit) : 0u;
                m1 |= e1 ? (1u << bit) : 0u;
            }
            n0.pp_h = n0.p_h, n0.p_h = h0, n0.p_s = s0;
            n1.pp_h = n1.p_h, n1.p_h = h1, n1.p_s = s1;
        };
        // bit b set: the row above finished row (y_base + b) is a peak
        auto flush = [&](int y_base) {
            if (m0 | m1) {
                for (unsigned m = m0; m; m &= m - 1) emit_peak(p, frame, part, y_base + __ffs(m) - 2, x0);
                for (unsigned m = m1; m; m &= m - 1) emit_peak(p, frame, part, y_base + __ffs(m) - 2, x1);
                m0 = m1 = 0;
            }
        };
        auto ld0 = [&](int i) { return have0 ? col0[(i - ilo) * RW] : NINF; };
        auto ld1 = [&](int i) { return have1 ? col1[(i - ilo) * RW] : NINF; };

        // window of the two columns: rows i-NB .. i+NB of the row-pass result (w?[NB] = row i).  Rows outside the
        // image hold 0 and are never selected under REFLECT_101 (make_win) / stand as 0 under a zero border.
        constexpr int NW = 2 * NB + 1;
        float w0[NW], w1[NW];
        auto ldr0 = [&](int i) { return (i >= 0 && i < h) ? ld0(i) : 0.f; };
        auto ldr1 = [&](int i) { return (i >= 0 && i < h) ? ld1(i) : 0.f; };
        auto load_window = [&](int i) { // all rows but the last (i + NB), which every use fetches itself
#pragma unroll
            for (int d = 0; d < NW - 1; ++d) w0[d] = ldr0(i - NB + d), w1[d] = ldr1(i - NB + d);
        };
        auto shift_window = [&]() {
#pragma unroll
            for (int d = 0; d < NW - 1; ++d) w0[d] = w0[d + 1], w1[d] = w1[d + 1];
        };
        auto smooth_rows = [&](int i, float (&s0)[S], float (&s1)[S]) { // the S image rows of feature row i, both columns
            w0[NW - 1] = ldr0(i + NB), w1[NW - 1] = ldr1(i + NB);
            // (a select-free copy of this for interior rows was tried: no faster, and twice the code to fetch)
            const Win9 v0 = make_win<NB, BZ>(w0[0], w0[NB - 1], w0[NB], w0[NB + 1], w0[NW - 1], i, h);
            const Win9 v1 = make_win<NB, BZ>(w1[0], w1[NB - 1], w1[NB], w1[NB + 1], w1[NW - 1], i, h);
            col_all2<S, R>(p.taps, v0, v1, s0, s1);
        };
        const int i_first = ia > 0 ? ia - 1 : ia;
        load_window(i_first);
        float s0[S], s1[S];
        const unsigned long long blk = s_blk[warp]; // bit r: block (this column group, feature row ilo + r) is active
        // the mask is the same in every lane; the vote makes that visible to the compiler (uniform branch,
        // so the shuffles below stay in convergent code)
        auto active = [&](int i) { return __any_sync(0xffffffffu, ((blk >> (i - ilo)) & 1ull) != 0) != 0; };
        bool open = false; // the last computed image row still waits for its decision
        auto close_block = [&](int y_next) { // the rows below are provably <= thresh: they count as -inf
            step(NINF, NINF, 0, true);
            flush(y_next);
            n0.pp_h = n0.p_h = n0.p_s = NINF, n1.pp_h = n1.p_h = n1.p_s = NINF;
            open = false;
        };
        bool stale = false; // the window does not hold the rows around i
        if (ia > 0) { // halo row above the tile: only its last image row is needed
            const int i = ia - 1;
            if (active(i)) {
                smooth_rows(i, s0, s1);
                step(s0[S - 1], s1[S - 1], 0, false);
                n0.p_s = NINF, n1.p_s = NINF; // that row belongs to the tile above
                shift_window();
            } else {
                stale = true;
            }
        }
        // Most feature rows of a column group are inactive on real maps (17 % active on the bench's): a run of them is
        // stepped over in one go - no loads, no per-row test - and the window is re-read at the next active row.
        for (int i = ia; i < ib; ++i) {
            if (!active(i)) {
                if (open) close_block(S * i);
                stale = true;
                if (STORE) { // the store stream keeps its row-by-row pace
                    store_units(3);
                    continue;
                }
                const unsigned long long rest = blk >> (i - ilo); // bit 0 = this (inactive) row
                const int run = rest ? __ffsll((long long)rest) - 1 : 64;
                i += min(run, ib - i) - 1;
                continue;
            }
            if (STORE) store_units(3);
            if (stale) {
                load_window(i);
                stale = false;
            }
            smooth_rows(i, s0, s1);
            // A feature row whose 64 x S smoothed pixels all stay <= thresh holds no peak and cannot outrank one: it
            // counts as -inf for the rows around it, exactly like a block the activity analysis ruled out - but decided
            // on the smoothed values themselves.  The 3x3 max (shuffles, compares, masks: half of this loop's
            // instructions, all on the ALU pipe) is then skipped; on maps with a noise floor that is most rows.
            float mx = fmaxf(s0[0], s1[0]);
#pragma unroll
            for (int ph = 1; ph < S; ++ph) mx = fmaxf(mx, fmaxf(s0[ph], s1[ph]));
            if (__any_sync(0xffffffffu, mx > p.thresh)) {
#pragma unroll
                for (int ph = 0; ph < S; ++ph) step(s0[ph], s1[ph], ph, true);
                flush(S * i);
                open = true;
            } else if (open) {
                close_block(S * i);
            }
            shift_window();
        }
        if (ib < h && active(ib)) { // halo row below the tile: its first image row closes the last row of the tile
            const int i = ib;
            if (stale) load_window(i);
            smooth_rows(i, s0, s1);
            step(s0[0], s1[0], 0, true);
            flush(S * i);
        } else if (open) { // bottom image edge, or an inactive block below: that row counts as -inf
            close_block(S * ib);
        }
    }

    if (STORE) store_units(st_total);
    stamp2(p, 4);
    // No tail: the unordered key lists are put into the reference's raster order by the limb kernel's CTAs, each
    // for its own two parts (order_part_peaks).  A per-frame "last tile" epilogue here kept every CTA - and its shared
    // memory - waiting a microsecond for the answer of its done-counter atomic (a third of all stall samples).
}

// ------------------------------------------------------------------------------------------------
// K2 generic path: any kernel size, any (non-integer) scale.  Reads the materialised up-sampled heat
// map tile + halo into shared memory with REFLECT_101 indexing, row pass, column pass, 3x3 max.
// ------------------------------------------------------------------------------------------------
#define G_TX 64
#define G_TY 32

// Row pass of cv::GaussianBlur's general RowFilter form over a staged region, register-blocked: a thread produces Q
// neighbouring outputs of one row from a sliding window it keeps in registers (one shared-memory load per tap and Q
// multiply-adds, instead of one load per multiply-add).  out[r][c] = k[0] in[r][c] + k[1] in[r][c+1] + ... left to
// right, the same products and sums as the one-output-per-thread form.
template <bool FIRST, int Q>
__device__ __forceinline__ void row_chunk(const float *__restrict__ taps, int j0, int K, float (&win)[2 * Q], float (&acc)[Q])
{
#pragma unroll
    for (int jj = 0; jj < Q; ++jj) {
        if (j0 + jj < K) {
            const float t = taps[j0 + jj];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float pr = __fmul_rn(t, win[jj + q]);
                acc[q] = (FIRST && jj == 0) ? pr : __fadd_rn(acc[q], pr);
            }
        }
    }
}

// Rows [r_a, r_b] and output columns [c_a, c_b] only (the part of the tile near values above the skip threshold).
__device__ __forceinline__ void row_pass_blocked(const Span<float> in, int IWs, const Span<float> out, int OW, const float *__restrict__ taps, int K, int r_a, int r_b,
                                                 int c_a, int c_b)
{
    constexpr int Q = 7;
    const int nblk = (c_b - c_a + Q) / Q, rows = r_b - r_a + 1;
    for (int it = threadIdx.x; it < rows * nblk; it += blockDim.x) {
        const int r = r_a + it / nblk, c0 = c_a + (it % nblk) * Q;
        const Span<float> src = in.from(r * IWs + c0);
        const int lim = IWs - c0; // entries of this row from src on (the last block reads past the outputs it keeps)
        float win[2 * Q], acc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) win[q] = q < lim ? src[q] : 0.f, acc[q] = 0.f;
        for (int j0 = 0; j0 < K; j0 += Q) {
#pragma unroll
            for (int q = 0; q < Q; ++q) win[Q + q] = j0 + Q + q < lim ? src[j0 + Q + q] : 0.f;
            if (j0 == 0) row_chunk<true, Q>(taps, j0, K, win, acc);
            else row_chunk<false, Q>(taps, j0, K, win, acc);
#pragma unroll
            for (int q = 0; q < Q; ++q) win[q] = win[Q + q];
        }
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (c0 + q <= c_b) out[r * OW + c0 + q] = acc[q];
    }
}

// Column pass (SymmColumnFilter: s = k[R] x[0]; s += k[R+j] (x[+j] + x[-j]), j = 1..R) for QR rows of one column per
// thread: the windows of rows above and below slide through registers (two loads per tap for QR outputs).
// x[i] = tmp[(i) * TW + c]; output row r reads tmp rows r .. r + 2R (its centre is row r + R).
template <int QR>
__device__ __forceinline__ void col_pass_blocked(const Span<float> tmp, int TW, int OR, const Span<float> out, const float *__restrict__ taps, int R, int y_first,
                                                 int x_first, int H, int W, int r_a, int r_b, int c_a, int c_b)
{
    const float NINF = -CUDART_INF_F;
    const int nblk = (r_b - r_a + QR) / QR, ncols = c_b - c_a + 1;
    for (int it = threadIdx.x; it < ncols * nblk; it += blockDim.x) {
        const int blk = it / ncols, c = c_a + (it - blk * ncols), r0 = r_a + blk * QR;
        const Span<float> col = tmp.from(c);
        const int last = OR + 2 * R - 1; // last row of tmp
        auto X = [&](int i) { return col[min(i, last) * TW]; }; // rows past the end feed outputs that are dropped
        float up[QR], dn[QR], sacc[QR];
        const float tR = taps[R];
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const float ctr = X(r0 + R + q);
            up[q] = dn[q] = ctr, sacc[q] = __fmul_rn(tR, ctr);
        }
        for (int j0 = 1; j0 <= R; j0 += QR) {
#pragma unroll
            for (int jj = 0; jj < QR; ++jj) {
                const int j = j0 + jj;
                if (j <= R) {
                    // after this step the up window holds x[r0+R+q+j], the down window x[r0+R+q-j]; logical entry q lives in
                    // ring slot (q + jj + 1) % QR resp. (q - jj - 1) mod QR, back in place after QR steps
                    up[jj % QR] = X(r0 + R + QR - 1 + j);
                    dn[(QR - 1 - jj) % QR] = X(r0 + R - j);
                    const float t = taps[R + j];
#pragma unroll
                    for (int q = 0; q < QR; ++q) {
                        const float u = up[(q + jj + 1) % QR], d = dn[((q - jj - 1) % QR + QR) % QR];
                        sacc[q] = __fadd_rn(sacc[q], __fmul_rn(t, __fadd_rn(u, d)));
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const int r = r0 + q;
            if (r <= r_b) {
                const int y = y_first + r, x = x_first + c;
                out[r * TW + c] = (y >= 0 && y < H && x >= 0 && x < W) ? sacc[q] : NINF;
            }
        }
    }
}

// Any scale (cv::resize INTER_AREA up-sampling = 2-tap linear with area-mode coefficients, src/post-process.h:24-49) and
// any kernel size: the tile of the up-sampled heat map (+ NMS and filter halo) is built in shared memory straight from
// the FEATURE map - horizontal 2-tap pass on the feature rows under the tile, vertical 2-tap blend per image row: the same
// float operations cv::resize performs through its row buffers - so no up-sampled map is read from HBM (or, in
// skeleton-only mode, ever written).  With p.conf_up != nullptr the tile is read from that materialised map instead.
__global__ void __launch_bounds__(OPP_THREADS) k2_peaks_generic(const K2Params p)
{
    extern __shared__ __align__(16) float smem[];
    const int frame = blockIdx.z, part = blockIdx.y;
    const int H = p.g.H, W = p.g.W, K = p.g.K, R = p.g.R, h = p.g.h, w = p.g.w;
    const int tiles_x = (W + G_TX - 1) / G_TX;
    const int x0 = (blockIdx.x % tiles_x) * G_TX, y0 = (blockIdx.x / tiles_x) * G_TY;
    const int IW = G_TX + 2 + 2 * R, IH = G_TY + 2 + 2 * R; // input region: outputs + NMS halo + filter halo
    const int TW = G_TX + 2;
    const Span<float> in = SPAN(float, smem, IH * IW);                     // [IH][IW]
    const Span<float> tmp = SPAN(float, smem + IH * IW, IH * TW);          // [IH][TW]  row-pass result
    const Span<float> sm = SPAN(float, tmp.p + IH * TW, (G_TY + 2) * TW);  // [G_TY+2][TW] smoothed
    // [nfr][IW] horizontally resized feature rows (dead before tmp is written; the launcher sizes the bytes behind `in` for it)
    const Span<float> T = SPAN(float, tmp.p, (size_t)min((long)h, ((long)IH * h + H - 1) / H + 3) * IW);
    // Only pixels inside the image are smoothed; the filter halo reflects, the NMS halo outside the
    // image is -inf.  Indices of halo pixels whose centre is outside the image are reflected too
    // (harmless: those smoothed values are discarded).
    // rows [r_a, r_b] x columns [c_a, c_b] of the smoothed tile (sm coordinates: row r is image row y0 - 1 + r) are computed;
    // the rest of it stands as -inf
    int r_a = 0, r_b = G_TY + 1, c_a = 0, c_b = TW - 1;
    if (p.conf_up) {
        int hot = 0;
        const float *plane = p.conf_up + ((size_t)frame * OPP_N_HEAT + part) * H * W;
        for (int t = threadIdx.x; t < IH * IW; t += blockDim.x) {
            const int yy = y0 - 1 - R + t / IW, xx = x0 - 1 - R + t % IW;
            int ry = reflect101(yy, H), rx = reflect101(xx, W);
            ry = clip_idx(ry, H), rx = clip_idx(rx, W);
            float v = __ldcg(plane + (size_t)ry * W + rx);
            // Python-path variant: 'SAME' zero padding of tf.nn.depthwise_conv2d (post_process.py:25-26)
            if (p.border_zero && (yy < 0 || yy >= H || xx < 0 || xx >= W)) v = 0.f;
            in[t] = v;
            hot |= v > p.skip_thresh;
        }
        // Same exact early-out as the fast kernel: the staged region is everything the tile's smoothed
        // pixels depend on; if all of it is <= thresh * (1 - 2^-13) no pixel of the tile can be a peak.
        if (!__syncthreads_or(hot)) return;
    } else {
        const float *plane = p.conf + ((size_t)frame * OPP_N_HEAT + part) * h * w;
        // image rows the tile's in-image pixels can reach, and the feature rows under them
        const int ylo = max(y0 - 1 - R, 0), yhi = min(y0 + G_TY + R, H - 1);
        const int f_lo = clip_idx(p.g.yofs[ylo], h), f_hi = clip_idx(p.g.yofs[yhi] + 1, h);
        const int nfr = f_hi - f_lo + 1;
        {
            // Which part of the tile can hold a peak at all?  Every pixel of the up-sampled map is a blend, with non-negative
            // weights that sum to 1 within 2^-23 per pass, of the (at most) 2 x 2 feature cells under it: pixels all of whose
            // cells are <= thresh (1 - 2^-13) stay below that bound within 2^-21, smooth to <= thresh, cannot be peaks and
            // cannot outrank one.  The bounding box of the pixels that cells ABOVE the bound can reach, grown by the filter
            // radius, is the only part of the tile worth smoothing; a tile with no such cell is done.
            __shared__ int s_bb[4];
            if (threadIdx.x == 0) s_bb[0] = s_bb[2] = 0x7fffffff, s_bb[1] = s_bb[3] = -1;
            __syncthreads();
            const int xlo = max(x0 - 1 - R, 0), xhi = min(x0 + G_TX + R, W - 1);
            const int c_lo = clip_idx(p.g.xofs[xlo], w), c_hi = clip_idx(p.g.xofs[xhi] + 1, w), ncf = c_hi - c_lo + 1;
            for (int t = threadIdx.x; t < nfr * ncf; t += blockDim.x) {
                const int fr = f_lo + t / ncf, fc = c_lo + t % ncf;
                if (plane[fr * w + fc] > p.skip_thresh) {
                    // pixels whose two taps include this cell: source index fr - 1 or fr, i.e. y h / H in [fr - 1, fr + 1);
                    // one pixel of slack on either side for the rounding of cv::resize's double-precision scale
                    const int ya = (int)(((long)(fr - 1) * H) / h) - 1, yb = (int)(((long)(fr + 1) * H + h - 1) / h) + 1;
                    const int xa = (int)(((long)(fc - 1) * W) / w) - 1, xb = (int)(((long)(fc + 1) * W + w - 1) / w) + 1;
                    atomicMin(&s_bb[0], ya), atomicMax(&s_bb[1], yb), atomicMin(&s_bb[2], xa), atomicMax(&s_bb[3], xb);
                }
            }
            __syncthreads();
            if (s_bb[1] < 0) return;
            r_a = max(s_bb[0] - R - (y0 - 1), 0), r_b = min(s_bb[1] + R - (y0 - 1), G_TY + 1);
            c_a = max(s_bb[2] - R - (x0 - 1), 0), c_b = min(s_bb[3] + R - (x0 - 1), TW - 1);
            if (r_a > r_b || c_a > c_b) return;
        }
        // region rows r_a .. r_b + 2R and region columns c_a .. c_b + 2R are what the two passes read
        const int t_a = c_a, t_n = c_b + 2 * R - c_a + 1;
        for (int t = threadIdx.x; t < nfr * t_n; t += blockDim.x) {
            const int fr = t / t_n, tc = t_a + (t - fr * t_n), xx = x0 - 1 - R + tc;
            const int rx = clip_idx(reflect101(xx, W), W);
            const float *S0 = plane + (f_lo + fr) * w;
            const int sx = p.g.xofs[rx];
            float v;
            if (rx < p.g.xmax) v = __fadd_rn(__fmul_rn(S0[sx], p.g.alpha[2 * rx]), __fmul_rn(S0[sx + 1], p.g.alpha[2 * rx + 1]));
            else v = __fmul_rn(S0[sx], 1.f);
            T[fr * IW + tc] = v;
        }
        __syncthreads();
        // vertical blend: a warp per region row (row index, source rows and coefficients once per row), lanes across it
        for (int i = r_a + (threadIdx.x >> 5); i <= r_b + 2 * R; i += blockDim.x >> 5) {
            const int yy = y0 - 1 - R + i;
            const bool row_out = yy < 0 || yy >= H;
            const int ry = min(max(clip_idx(reflect101(yy, H), H), ylo), yhi); // rows beyond feed dropped outputs only
            const int sy = p.g.yofs[ry];
            const Span<float> T0 = T.from((clip_idx(sy, h) - f_lo) * IW), T1 = T.from((clip_idx(sy + 1, h) - f_lo) * IW);
            const float b0 = p.g.beta[2 * ry], b1 = p.g.beta[2 * ry + 1];
            for (int tc = t_a + (threadIdx.x & 31); tc < t_a + t_n; tc += 32) {
                const int xx = x0 - 1 - R + tc;
                float v = __fadd_rn(__fmul_rn(T0[tc], b0), __fmul_rn(T1[tc], b1));
                if (p.border_zero && (row_out || xx < 0 || xx >= W)) v = 0.f;
                in[i * IW + tc] = v;
            }
        }
        __syncthreads();
    }
    const float NINF_ = -CUDART_INF_F;
    const bool whole = r_a == 0 && r_b == G_TY + 1 && c_a == 0 && c_b == TW - 1;
    if (!whole)
        for (int t = threadIdx.x; t < (G_TY + 2) * TW; t += blockDim.x) sm[t] = NINF_;
    if ((K == 3 || K == 5) && !p.border_zero) { // OpenCV's symmetric-small row forms
        const int ncols = c_b - c_a + 1;
        for (int t = threadIdx.x; t < (r_b + 2 * R - r_a + 1) * ncols; t += blockDim.x) {
            const int r = r_a + t / ncols, c = c_a + t % ncols;
            const Span<float> q = in.from(r * IW + c);
            float sv = __fadd_rn(__fmul_rn(q[R], p.taps[R]), __fmul_rn(__fadd_rn(q[R - 1], q[R + 1]), p.taps[R + 1]));
            if (K == 5) sv = __fadd_rn(sv, __fmul_rn(__fadd_rn(q[R - 2], q[R + 2]), p.taps[R + 2]));
            tmp[r * TW + c] = sv;
        }
    } else {
        row_pass_blocked(in, IW, tmp, TW, p.taps, K, r_a, r_b + 2 * R, c_a, c_b);
    }
    __syncthreads();
    col_pass_blocked<6>(tmp, TW, G_TY + 2, sm, p.taps, R, y0 - 1, x0 - 1, H, W, r_a, r_b, c_a, c_b);
    __syncthreads();
    for (int t = threadIdx.x; t < G_TY * G_TX; t += blockDim.x) {
        const int r = t / G_TX, c = t % G_TX;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        const int o = (r + 1) * TW + (c + 1);
        const float s = sm[o];
        if (!(s > p.thresh)) continue;
        float m = fmaxf(fmaxf(sm[o - TW - 1], sm[o - TW]), sm[o - TW + 1]);
        m = fmaxf(m, fmaxf(fmaxf(sm[o - 1], s), sm[o + 1]));
        m = fmaxf(m, fmaxf(fmaxf(sm[o + TW - 1], sm[o + TW]), sm[o + TW + 1]));
        if (s == m) emit_peak(p, frame, part, y, x);
    }
}

// ------------------------------------------------------------------------------------------------
// K2 generic path at an INTEGER scale (any kernel size <= 63, either border rule): the up-sampled map is the feature
// map replicated S x S, so (a) the tile is staged straight from the feature maps - no materialised heat map is read -
// and (b) the row pass, the expensive half for large kernels, is evaluated once per distinct FEATURE row under the
// tile instead of once per image row (S times fewer).  The column pass then looks the image rows up through a small
// row map (REFLECT_101 or zero border -> row of the row-pass result, or "zero").  Every smoothed value is produced by
// the same IEEE operations on the same operands, in the same order, as k2_peaks_generic: results are identical.
// Used for k = 19..63 at x8 (k = 25 is the fixed size of the reference's Python graph) and for the Python variant.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OPP_THREADS) k2_peaks_generic_rep(const K2Params p)
{
    extern __shared__ __align__(16) float smem[];
    const int frame = blockIdx.z, part = blockIdx.y;
    const int H = p.g.H, W = p.g.W, K = p.g.K, R = p.g.R, S = p.g.S, h = p.g.h, w = p.g.w;
    const int tiles_x = (W + G_TX - 1) / G_TX;
    const int x0 = (blockIdx.x % tiles_x) * G_TX, y0 = (blockIdx.x / tiles_x) * G_TY;
    const int IW = G_TX + 2 + 2 * R, IH = G_TY + 2 + 2 * R; // image region: outputs + NMS halo + filter halo
    const int TW = G_TX + 2;
    const bool zero = p.border_zero != 0;
    // distinct feature rows under the (border-mapped) image rows of the region
    const int ylo = max(y0 - 1 - R, 0), yhi = min(y0 + G_TY + R, H - 1);
    const int f_lo = ylo / S, nfr = yhi / S - f_lo + 1;
    float *in = smem;               // [nfr][IW]   feature rows, columns already expanded and border-mapped
    float *tmp = in + nfr * IW;     // [nfr+1][TW] row-pass result; row nfr is all zeros (rows beyond a zero border)
    float *sm = tmp + (nfr + 1) * TW; // [G_TY+2][TW] smoothed
    int *rowmap = reinterpret_cast<int *>(sm + (G_TY + 2) * TW); // [IH] image row of the region -> offset of its row in tmp
    const float *plane = p.conf + ((size_t)frame * OPP_N_HEAT + part) * h * w;
    for (int i = threadIdx.x; i < IH; i += blockDim.x) {
        const int yy = y0 - 1 - R + i;
        int m;
        if (zero) {
            m = (yy < 0 || yy >= H) ? nfr : yy / S - f_lo;
        } else {
            const int ry = clip_idx(reflect101(yy, H), H); // rows whose centre is outside the image are never used
            m = min(max(ry / S - f_lo, 0), nfr - 1);
        }
        rowmap[i] = m * TW;
    }
    for (int c = threadIdx.x; c < TW; c += blockDim.x) tmp[nfr * TW + c] = 0.f;
    int hot = 0;
    for (int t = threadIdx.x; t < nfr * IW; t += blockDim.x) {
        const int fr = t / IW, xx = x0 - 1 - R + t % IW;
        float v;
        if (zero && (xx < 0 || xx >= W)) {
            v = 0.f;
        } else {
            const int rx = clip_idx(reflect101(xx, W), W);
            v = __ldcg(plane + (size_t)(f_lo + fr) * w + rx / S);
        }
        in[t] = v;
        hot |= v > p.skip_thresh;
    }
    // same exact early-out as the other peak kernels: the staged rows are everything the tile's pixels depend on
    if (!__syncthreads_or(hot)) {
        return;
    }
    for (int t = threadIdx.x; t < nfr * TW; t += blockDim.x) {
        const int r = t / TW, c = t % TW;
        const float *q = in + r * IW + c; // q[j] = tap j of output column x0-1+c
        float s;
        if (K == 3 && !zero) {
            s = __fadd_rn(__fmul_rn(q[R], p.taps[R]), __fmul_rn(__fadd_rn(q[R - 1], q[R + 1]), p.taps[R + 1]));
        } else if (K == 5 && !zero) {
            s = __fadd_rn(__fmul_rn(q[R], p.taps[R]), __fmul_rn(__fadd_rn(q[R - 1], q[R + 1]), p.taps[R + 1]));
            s = __fadd_rn(s, __fmul_rn(__fadd_rn(q[R - 2], q[R + 2]), p.taps[R + 2]));
        } else {
            s = __fmul_rn(p.taps[0], q[0]);
            for (int j = 1; j < K; ++j) s = __fadd_rn(s, __fmul_rn(p.taps[j], q[j]));
        }
        tmp[t] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (G_TY + 2) * TW; t += blockDim.x) {
        const int r = t / TW, c = t % TW;
        const int y = y0 - 1 + r, x = x0 - 1 + c;
        float s = -CUDART_INF_F;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const int *rm = rowmap + r + R; // rm[j] = offset in tmp of the row standing for image row y + j
            const float *tc = tmp + c;
            s = __fmul_rn(p.taps[R], tc[rm[0]]);
            for (int j = 1; j <= R; ++j) s = __fadd_rn(s, __fmul_rn(p.taps[R + j], __fadd_rn(tc[rm[j]], tc[rm[-j]])));
        }
        sm[t] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < G_TY * G_TX; t += blockDim.x) {
        const int r = t / G_TX, c = t % G_TX;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        const float *q = sm + (r + 1) * TW + (c + 1);
        const float s = q[0];
        if (!(s > p.thresh)) continue;
        float m = fmaxf(fmaxf(q[-TW - 1], q[-TW]), q[-TW + 1]);
        m = fmaxf(m, fmaxf(fmaxf(q[-1], q[0]), q[1]));
        m = fmaxf(m, fmaxf(fmaxf(q[TW - 1], q[TW]), q[TW + 1]));
        if (s == m) emit_peak(p, frame, part, y, x);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: one CTA per (limb, frame): score all (a, b) peak pairs, order-preserving compaction of the
// accepted candidates, std::sort-order sort, greedy matching.  The last limb of a frame to finish
// assembles the frame's humans.
// ------------------------------------------------------------------------------------------------
// roundpaf (src/paf.cpp:337): (int)(v + 0.5) with the add in double.  For 0 <= v < 2^23 the double sum
// t + f + 0.5 (t = floor v, f = v - t, both exact in float) truncates to t + (f >= 0.5): no FP64 needed.
__device__ __forceinline__ int round_paf(float v)
{
    if (v >= 0.f && v < 8388608.f) {
        const int t = (int)v;
        return t + (__fsub_rn(v, (float)t) >= 0.5f ? 1 : 0);
    }
    return (int)__dadd_rn((double)v, 0.5);
}

struct Cand {
    int i1, i2;
    float s;
};

// PAF line integral of one (a, b) peak pair: src/paf.cpp:86-127 with get_paf_vectors :313-335.  QUICK evaluates
// only the four middle samples (i = 3..6) and answers "can this pair still be accepted?": acceptance needs more
// than 8 of the 10 samples above THRESH_VECTOR_SCORE (:119-121), so two failures among ANY samples rule the pair
// out - exactly, whatever the others say.  Most pairs of a frame join peaks of different people, whose segment
// crosses empty PAF in the middle: they leave after 4 samples instead of 10 and the survivors (re-evaluated in
// full, same operations in the same order) fill whole warps.
struct PairCtx {
    Span<const float> px, py; // the limb's x / y PAF planes (feature resolution)
    const OppGeom *g;
    Span<const float> steps;   // [max(H, W)] d -> (float)d / 10.f (IEEE), or null: the STEP_X / STEP_Y divisions as a table
    Span<const unsigned> weak; // bit per feature cell: |px| + |py| so small that no sample there can exceed THRESH_VECTOR_SCORE (or null)
    int w, H, sshift;
    float thr;
};

// roundpaf for 0 <= v < 2^22 without conversion instructions (F2I / I2F run on the quarter-rate XU pipe, and the pair
// filter rounds eight coordinates per pair): v + 2^23 rounds v to the nearest integer, ties to even; floor(v + 0.5) -
// what (int)((double)v + 0.5) is for v >= 0 - differs from that only at a tie that went down, where v - r = +0.5.
__device__ __forceinline__ int round_paf_small(float v)
{
    const float r = __fadd_rn(v, 8388608.f);
    const float d = __fsub_rn(v, __fsub_rn(r, 8388608.f)); // exact
    return (__float_as_int(r) - 0x4B000000) + (d >= 0.5f ? 1 : 0);
}

// First stage of the pair filter at power-of-two scales: the four middle samples' CELLS only.  A sample's score is
// v . paf(cell) with |v.x|, |v.y| <= 1 + 2^-23, so in a cell with |px| + |py| <= thresh (1 - 2^-13) it cannot exceed the
// threshold whatever the direction: two such samples rule the pair out (cnt <= 8).  No square root, no division, no
// PAF load - and most pairs of a crowded frame join peaks of different people across empty PAF.
__device__ __forceinline__ bool pair_may_pass(const PairCtx &c, const int2 A, const int2 B)
{
    const int dx = B.x - A.x, dy = B.y - A.y;
    if ((dx | dy) == 0) return false;
    const float step_x = c.steps.p ? copysignf(c.steps[abs(dx)], (float)dx) : __fdiv_rn((float)dx, 10.f);
    const float step_y = c.steps.p ? copysignf(c.steps[abs(dy)], (float)dy) : __fdiv_rn((float)dy, 10.f);
    const float ax = (float)A.x, ay = (float)A.y;
    int weak = 0;
#pragma unroll
    for (int i = 3; i < 7; ++i) {
        const int lx = round_paf_small(__fadd_rn(ax, __fmul_rn((float)i, step_x)));
        const int ly = round_paf_small(__fadd_rn(ay, __fmul_rn((float)i, step_y)));
        const int cell = (ly >> c.sshift) * c.w + (lx >> c.sshift);
        weak += (c.weak[cell >> 5] >> (cell & 31)) & 1u;
    }
    return weak < 2;
}

template <bool QUICK>
__device__ __forceinline__ bool score_pair(const PairCtx &c, const int2 A, const int2 B, float &crit2)
{
    const int dx = B.x - A.x, dy = B.y - A.y;
    // norm = (float)sqrt((double)l2)  (src/paf.cpp:91).  Rounding sqrt to 53 and then to 24 bits equals
    // rounding it to 24 bits directly (double rounding is innocuous for sqrt when 53 >= 2*24 + 2),
    // so the single-precision IEEE sqrt gives the same float whenever l2 is exact in float.
    const int l2 = dx * dx + dy * dy;
    if (l2 == 0) return false; // `norm < 1e-12` is true only for coincident peaks
    const float norm = l2 < (1 << 24) ? __fsqrt_rn((float)l2) : (float)sqrt((double)l2);
    const float vx = __fdiv_rn((float)dx, norm), vy = __fdiv_rn((float)dy, norm);
    // STEP_X, STEP_Y = d / 10.f  (:321-322); division rounds symmetrically, so the table of |d| serves both signs
    const float step_x = c.steps.p ? copysignf(c.steps[abs(dx)], (float)dx) : __fdiv_rn((float)dx, 10.f);
    const float step_y = c.steps.p ? copysignf(c.steps[abs(dy)], (float)dy) : __fdiv_rn((float)dy, 10.f);
    float scores = 0.f;
    int cnt = 0;
    // Not unrolled all the way: the limb kernel runs most of its code once or a few times per CTA, i.e. cold, and
    // straight-line code costs more in instruction fetch than the loop costs in issue slots (10 samples unrolled: +1.5 us
    // per limb CTA on the one-frame latency path, -7 % throughput on crowded batches).
#pragma unroll 2
    for (int i = QUICK ? 3 : 0; i < (QUICK ? 7 : 10); ++i) {
        // roundpaf(peak1.x + i * STEP_X): float mul, float add, double +0.5, truncate  :325-326,337
        const float fx = __fadd_rn((float)A.x, __fmul_rn((float)i, step_x));
        const float fy = __fadd_rn((float)A.y, __fmul_rn((float)i, step_y));
        // No clamp (the reference has none): |i * STEP| <= 0.9 |d| (1 + 2^-22) and |d| >= 1 on a moving axis, so every
        // sample rounds to a pixel between the two peaks, which lie inside the image.
        const int lx = round_paf(fx), ly = round_paf(fy);
        float vpx, vpy;
        if (c.sshift >= 0) { // replication by a power of two: one shared index, no integer division
            const int fi = (ly >> c.sshift) * c.w + (lx >> c.sshift);
            vpx = c.px[fi], vpy = c.py[fi];
        } else {
            vpx = upsample_at(*c.g, c.px.p, ly, lx);
            vpy = upsample_at(*c.g, c.py.p, ly, lx);
        }
        const float score = __fadd_rn(__fmul_rn(vx, vpx), __fmul_rn(vy, vpy)); // :108-109
        scores = __fadd_rn(scores, score);
        cnt += (score > c.thr);
    }
    if (QUICK) return cnt >= 3;
    // scores / STEP_PAF + std::min(0.0, 0.5 * height / norm - 1.0)   :115-116
    const float s10 = __fdiv_rn(scores, 10.f);
    if (norm <= 0.5f * (float)c.H) {
        crit2 = s10; // 0.5*H/norm >= 1 exactly, so the penalty is min(0.0, >= 0) = 0 and (float)((double)s10 + 0.0) == s10
    } else {
        const double pen = __dsub_rn(__ddiv_rn(0.5 * (double)c.H, (double)norm), 1.0);
        crit2 = (float)__dadd_rn((double)s10, pen < 0.0 ? pen : 0.0);
    }
    return cnt > 8 && crit2 > 0.f;
}

__device__ __forceinline__ bool cand_gt(const Cand &a, const Cand &b) { return a.s > b.s; }

// libstdc++ (GCC 13) std::sort with comp(a,b) = a.score > b.score, as called at src/paf.cpp:151-152.
// bits/stl_algo.h: __sort :1942, __introsort_loop :1918, __unguarded_partition_pivot :1893,
// __move_median_to_first :85, __unguarded_partition :1871, __final_insertion_sort :1854,
// __partial_sort :1905 (heap sort fallback).  Sequential: element movement decides tie order.
__device__ void sift_down(Cand *first, long hole, long len, Cand value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (cand_gt(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && cand_gt(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

__device__ void heap_sort_range(Cand *first, Cand *last)
{
    const long len = last - first;
    if (len >= 2) {
        for (long parent = (len - 2) / 2;; --parent) {
            const Cand v = first[parent];
            sift_down(first, parent, len, v);
            if (parent == 0) break;
        }
    }
    while (last - first > 1) {
        --last;
        const Cand v = *last;
        *last = *first;
        sift_down(first, 0, last - first, v);
    }
}

__device__ __forceinline__ void cswap(Cand *a, Cand *b)
{
    const Cand t = *a;
    *a = *b;
    *b = t;
}

__device__ void unguarded_linear_insert(Cand *last)
{
    const Cand val = *last;
    Cand *next = last - 1;
    while (cand_gt(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

__device__ void insertion_sort_range(Cand *first, Cand *last)
{
    if (first == last) return;
    for (Cand *i = first + 1; i != last; ++i) {
        if (cand_gt(*i, *first)) {
            const Cand val = *i;
            for (Cand *q = i; q != first; --q) *q = *(q - 1);
            *first = val;
        } else
            unguarded_linear_insert(i);
    }
}

__device__ __noinline__ void std_sort_desc(Cand *v, int n)
{
    if (n <= 0) return;
    struct Range {
        int first, last, depth;
    };
    Range stack[48];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    stack[sp++] = Range{0, n, 2 * lg};
    while (sp > 0) {
        const Range r = stack[--sp];
        Cand *first = v + r.first, *last = v + r.last;
        int depth = r.depth;
        while (last - first > 16) {
            if (depth == 0) {
                heap_sort_range(first, last);
                break;
            }
            --depth;
            Cand *a = first + 1, *b = first + (last - first) / 2, *c = last - 1;
            // __move_median_to_first(first, a, b, c)
            if (cand_gt(*a, *b)) {
                if (cand_gt(*b, *c))
                    cswap(first, b);
                else if (cand_gt(*a, *c))
                    cswap(first, c);
                else
                    cswap(first, a);
            } else if (cand_gt(*a, *c))
                cswap(first, a);
            else if (cand_gt(*b, *c))
                cswap(first, c);
            else
                cswap(first, b);
            // __unguarded_partition(first + 1, last, first)
            Cand *lo = first + 1, *hi = last;
            for (;;) {
                while (cand_gt(*lo, *first)) ++lo;
                --hi;
                while (cand_gt(*first, *hi)) --hi;
                if (!(lo < hi)) break;
                cswap(lo, hi);
                ++lo;
            }
            // recurse on [lo, last), continue with [first, lo): disjoint ranges, order is immaterial
            if (sp < 48) stack[sp++] = Range{(int)(lo - v), (int)(last - v), depth};
            last = lo;
        }
    }
    if (n > 16) {
        insertion_sort_range(v, v + 16);
        for (Cand *i = v + 16; i != v + n; ++i) unguarded_linear_insert(i);
    } else
        insertion_sort_range(v, v + n);
}

// The partition loop of that std::sort (everything before __final_insertion_sort) in parallel, whole CTA, in place.
// tests/test_sort_model.py is the executable CPU model and carries the argument:
//   * __unguarded_partition is a Hoare partition: with A = the positions (ascending) whose element does not beat the
//     pivot and B = the positions (descending) the pivot does not beat, the sequential loop swaps A[k] <-> B[k] for
//     every k with A[k] < B[k] and returns cut = A[K] if A[K] < B[K-1] else B[K-1] (K swaps; A[0] when K = 0).  A warp
//     builds both lists with ballots and does the swaps at once;
//   * the two sides of a cut are independent: the ranges of one level go to different warps;
//   * a range whose depth budget is spent is heap-sorted sequentially, as libstdc++ does.
// What remains (__final_insertion_sort) is an insertion sort = the stable order of the array left behind, which the
// caller produces with a rank sort.  pos_a / pos_b: scratch, n entries each; rng: scratch, 4 (n / 17 + 2) ints.
__device__ __noinline__ void std_sort_partition_rounds(const Span<Cand> v, int n, const Span<unsigned short> pos_a, const Span<unsigned short> pos_b, const Span<int> rng)
{
    __shared__ int s_rcnt[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int max_r = n / 17 + 2;
    const Span<int> rng_lo[2] = {rng, rng.from(2 * max_r)}; // per level: {first | last << 16, depth}
    if (tid == 0) {
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) ++lg;
        s_rcnt[0] = n > 16 ? 1 : 0, s_rcnt[1] = 0;
        rng[0] = 0 | (n << 16), rng[1] = 2 * lg;
    }
    __syncthreads();
    for (int cur = 0;; cur ^= 1) {
        const int ncur = s_rcnt[cur];
        if (ncur == 0) break;
        for (int r = warp; r < ncur; r += nwarps) {
            const int packed = rng_lo[cur][2 * r], depth = rng_lo[cur][2 * r + 1];
            const int f = packed & 0xffff, l = (packed >> 16) & 0xffff;
            if (depth == 0) {
                if (lane == 0) heap_sort_range(v.p + f, v.p + l);
                continue;
            }
            if (lane == 0) { // __move_median_to_first(first, first + 1, mid, last - 1)
                Cand *first = v.p + f, *a = first + 1, *b = first + (l - f) / 2, *c = v.p + l - 1;
                if (cand_gt(*a, *b)) {
                    if (cand_gt(*b, *c))
                        cswap(first, b);
                    else if (cand_gt(*a, *c))
                        cswap(first, c);
                    else
                        cswap(first, a);
                } else if (cand_gt(*a, *c))
                    cswap(first, a);
                else if (cand_gt(*b, *c))
                    cswap(first, c);
                else
                    cswap(first, b);
            }
            __syncwarp();
            const float piv = v[f].s;
            // both lists in one ascending pass (B is read backwards afterwards)
            int cnt_a = 0, cnt_b = 0;
            const unsigned lt = (1u << lane) - 1;
            for (int base = f + 1; base < l; base += 32) {
                const int q = base + lane;
                const bool in = q < l;
                const float sq = in ? v[q].s : 0.f;
                const bool fa = in && !(sq > piv), fb = in && !(piv > sq);
                const unsigned ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
                if (fa) pos_a[f + cnt_a + __popc(ma & lt)] = (unsigned short)q;
                if (fb) pos_b[f + cnt_b + __popc(mb & lt)] = (unsigned short)q;
                cnt_a += __popc(ma), cnt_b += __popc(mb);
            }
            __syncwarp();
            auto A = [&](int k) { return (int)pos_a[f + k]; };
            auto B = [&](int k) { return (int)pos_b[f + cnt_b - 1 - k]; };
            const int m = min(cnt_a, cnt_b);
            int K = 0; // A ascends and B descends: A[k] < B[k] holds for a prefix of k
            for (int base = 0; base < m; base += 32) {
                const int k = base + lane;
                const unsigned mk = __ballot_sync(0xffffffffu, k < m && A(k) < B(k));
                K += __popc(mk);
                if (mk != 0xffffffffu) break;
            }
            const int cut = K == 0 ? A(0) : ((K < cnt_a && A(K) < B(K - 1)) ? A(K) : B(K - 1));
            for (int k = lane; k < K; k += 32) cswap(&v[A(k)], &v[B(k)]);
            if (lane == 0) {
                if (l - cut > 16) {
                    const int i = atomicAdd(&s_rcnt[cur ^ 1], 1);
                    rng_lo[cur ^ 1][2 * i] = cut | (l << 16), rng_lo[cur ^ 1][2 * i + 1] = depth - 1;
                }
                if (cut - f > 16) {
                    const int i = atomicAdd(&s_rcnt[cur ^ 1], 1);
                    rng_lo[cur ^ 1][2 * i] = f | (cut << 16), rng_lo[cur ^ 1][2 * i + 1] = depth - 1;
                }
            }
        }
        __syncthreads(); // this level's swaps and child ranges are complete
        if (tid == 0) s_rcnt[cur] = 0;
        __syncthreads();
    }
}

// The candidates of one limb in std::sort's order (src/paf.cpp:151-152), whole CTA.  cand0: the n candidates (in a-major /
// b-minor order, or - `unordered` - in any order, identified by their pair index (i1 - ofs_a) * nb + (i2 - ofs_b));
// cand1: a second buffer of the same size; scratch: 4 (n / 17 + 2) ints.  Returns whichever buffer holds the result.
__device__ __forceinline__ Span<Cand> sort_candidates_desc(const Span<Cand> cand0, const Span<Cand> cand1, int n_cand, bool unordered, int ofs_a, int ofs_b,
                                                           unsigned nb_u, bool in_smem, const Span<int> scratch)
{
    const int tid = threadIdx.x;
    if (n_cand <= 1) return cand0;
    int tie = 0;
    if (n_cand <= 4096) {
        // rank by (score descending, position ascending): the order itself when no two scores are equal, and
        // the stable order = what __final_insertion_sort leaves when run after the partition rounds below
        auto rank_sort = [&]() {
            int any_tie = 0;
            for (int t = tid; t < n_cand; t += blockDim.x) {
                const float s = cand0[t].s;
                int rank = 0;
                for (int u = 0; u < n_cand; ++u) {
                    const float su = cand0[u].s;
                    rank += (su > s) | (su == s && u < t);
                    any_tie |= (su == s && u != t);
                }
                cand1[rank] = cand0[t];
            }
            return __syncthreads_or(any_tie);
        };
        tie = rank_sort();
        if (!tie) return cand1;
        if (unordered) {
            // std::sort's input is the candidate list in a-major / b-minor order (src/paf.cpp:93-131): put the
            // unordered list into that order first (rank by pair index, which is unique), back into cand0
            for (int t = tid; t < n_cand; t += blockDim.x) {
                const Cand ct = cand0[t];
                const unsigned key = (unsigned)(ct.i1 - ofs_a) * nb_u + (unsigned)(ct.i2 - ofs_b);
                int rank = 0;
                for (int u = 0; u < n_cand; ++u) rank += ((unsigned)(cand0[u].i1 - ofs_a) * nb_u + (unsigned)(cand0[u].i2 - ofs_b)) < key;
                cand1[rank] = ct;
            }
            __syncthreads();
            for (int t = tid; t < n_cand; t += blockDim.x) cand0[t] = cand1[t];
            __syncthreads();
        }
        if (in_smem && n_cand <= 0xffff) {
            // tied scores: std::sort's element movement decides.  cand1 serves as scratch in between.
            const Span<unsigned short> pos_a = SPAN(unsigned short, reinterpret_cast<unsigned short *>(cand1.p), n_cand);
            const Span<unsigned short> pos_b = SPAN(unsigned short, reinterpret_cast<unsigned short *>(cand1.p) + n_cand, n_cand);
            std_sort_partition_rounds(cand0, n_cand, pos_a, pos_b, scratch);
            rank_sort();
            return cand1;
        }
    }
    if (tid == 0) std_sort_desc(cand0.p, n_cand); // large lists, or lists in global memory: the sequential emulation
    __syncthreads();
    return cand0;
}

// Test entry: sorts n candidates (global memory, given order) the way the limb kernel does and writes them back in sorted
// order.  mode 0: parallel form (shared memory, n <= 4096), 1: sequential emulation.  One CTA.
__global__ void __launch_bounds__(OPP_THREADS) k_debug_sort(Cand *v, int n, int mode)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Span<Cand> c0 = SPAN(Cand, reinterpret_cast<Cand *>(smem_raw), n), c1 = SPAN(Cand, reinterpret_cast<Cand *>(smem_raw) + n, n);
    const Span<int> scratch = SPAN(int, reinterpret_cast<int *>(c1.p + n), 4 * (n / 17 + 2));
    Span<Cand> sorted = SPAN(Cand, v, n);
    if (mode == 0) {
        for (int t = threadIdx.x; t < n; t += blockDim.x) c0[t] = v[t];
        __syncthreads();
        sorted = sort_candidates_desc(c0, c1, n, false, 0, 0, 1u, true, scratch);
    } else {
        if (threadIdx.x == 0) std_sort_desc(v, n);
    }
    __syncthreads();
    if (mode == 0)
        for (int t = threadIdx.x; t < n; t += blockDim.x) v[t] = sorted[t];
}

__device__ __forceinline__ void stamp(const K3Params &p, int frame, int pair_id, int slot)
{
    if (p.times && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.times[((size_t)frame * OPP_N_PAIRS + pair_id) * 12 + slot] = t;
    }
}

// words of a partial human (human_ref_t, include/openpose-plus/human.h:57-77)
#define HR_WORDS 21
#define HR_ID 0
#define HR_PART 1
#define HR_SCORE 19
#define HR_NPARTS 20

// Person assembly for one frame (src/paf.cpp:177-262) + output (:292-310), run by the limb CTA that
// finished last.  Everything it needs (connection counts, all connections, peak x/y/score) is staged
// into shared memory in two parallel round trips; the order-dependent part then runs in one warp
// without touching global memory, and the output is written by the whole CTA.
__device__ void assemble_frame(const K3Params &p, int frame, unsigned char *smem_raw, const int *pofs /* shared: 19 part offsets */)
{
    const Span<int> hr = SPAN(int, reinterpret_cast<int *>(smem_raw + p.off_href), p.capH * HR_WORDS);                     // [capH][21]
    const Span<int2> s_pk = SPAN(int2, reinterpret_cast<int2 *>(smem_raw + p.off_score), p.pk_cap);                          // [n_peaks] {x | y << 16, score bits} (score_in_smem)
    const Span<opp_conn_t> s_conn = SPAN(opp_conn_t, reinterpret_cast<opp_conn_t *>(smem_raw + p.off_conn), p.conn_cap); // all connections of the frame, or one limb's
    const Span<int> s_keep = SPAN(int, reinterpret_cast<int *>(smem_raw + p.off_keep), p.capH);                             // [capH] surviving humans, output order
    __shared__ int s_state[8];                                                  // n, -, flags, merges, n_out
    __shared__ int s_nc[OPP_N_PAIRS], s_coff[OPP_N_PAIRS + 1];
    const int capH = p.capH, capP = p.capP;
    const int n_peaks = pofs[OPP_N_PARTS];
    const opp_peak_t *peaks = p.peaks + (size_t)frame * OPP_N_PARTS * capP;
    if (threadIdx.x <= OPP_N_PARTS) p.part_ofs[frame * (OPP_N_PARTS + 1) + threadIdx.x] = pofs[threadIdx.x]; // for opp_debug_fetch
    const bool pk_smem = p.score_in_smem != 0 && n_peaks <= p.pk_cap;
    if (threadIdx.x < 32) { // connection counts of the 19 limbs and their exclusive prefix sums, one warp scan
        const int ln = threadIdx.x;
        const int nc = ln < OPP_N_PAIRS ? __ldcg(p.n_conns + frame * OPP_N_PAIRS + ln) : 0;
        int incl = nc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (ln >= o) incl += v;
        }
        if (ln < OPP_N_PAIRS) s_nc[ln] = nc, s_coff[ln + 1] = incl;
        if (ln == 0) s_coff[0] = 0;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 40) s_state[threadIdx.x - 32] = 0;
    // Forest form of the 17 tree limbs (see below): s_c1[limb][local index of a first-part peak] = flattened index of
    // the limb's connection that starts at that peak (0xffff: none); s_in = bitmap over peak ids, set when the peak
    // enters a human as the SECOND part of a tree-limb connection; s_lmb[t] = limb of flattened connection t.
    const Span<unsigned short> s_c1 = SPAN(unsigned short, reinterpret_cast<unsigned short *>(smem_raw + p.off_owner), 17 * capP);                               // [17][capP]
    const Span<unsigned> s_in = SPAN(unsigned, reinterpret_cast<unsigned *>(smem_raw + p.off_owner + (((size_t)17 * capP * 2 + 15) & ~(size_t)15)), (OPP_N_PARTS * capP + 31) / 32); // [ceil(18 capP / 32)]
    const Span<unsigned char> s_lmb = SPAN(unsigned char, reinterpret_cast<unsigned char *>(s_in.p + (((OPP_N_PARTS * capP + 31) / 32 + 3) & ~3)), p.conn_cap); // [conn_cap]
    __shared__ int s_pofs[OPP_N_PARTS + 1];
    __shared__ int s_nwarp[OPP_THREADS / 32];
    __shared__ unsigned char s_pa[OPP_N_PAIRS], s_pb[OPP_N_PAIRS]; // c_pair_a / c_pair_b for lane-divergent limb indices
    if (threadIdx.x >= 96 && threadIdx.x < 96 + OPP_N_PAIRS) s_pa[threadIdx.x - 96] = (unsigned char)c_pair_a[threadIdx.x - 96], s_pb[threadIdx.x - 96] = (unsigned char)c_pair_b[threadIdx.x - 96];
    if (threadIdx.x >= 64 && threadIdx.x < 64 + OPP_N_PARTS + 1) s_pofs[threadIdx.x - 64] = pofs[threadIdx.x - 64];
    if (p.owner_in_smem) { // cleared before the frame's connection count is known
        for (int t = threadIdx.x; t < 17 * capP; t += blockDim.x) s_c1[t] = 0xffff;
        for (int t = threadIdx.x; t < (OPP_N_PARTS * capP + 31) / 32; t += blockDim.x) s_in[t] = 0u;
        for (int t = threadIdx.x; t < capH * OPP_N_PARTS; t += blockDim.x) hr[(t / OPP_N_PARTS) * HR_WORDS + HR_PART + t % OPP_N_PARTS] = -1;
    }
    if (pk_smem)
        for (int t = threadIdx.x; t < n_peaks; t += blockDim.x)
            s_pk[t] = make_int2(__ldcg(&peaks[t].x) | (__ldcg(&peaks[t].y) << 16), __float_as_int(__ldcg(&peaks[t].score)));
    __syncthreads();
    // frames with more connections than the staging area holds are assembled limb by limb (same code, one warp)
    const bool all_conns = p.conns_in_smem != 0 && s_coff[OPP_N_PAIRS] <= p.conn_cap;
    const bool use_owner = p.owner_in_smem != 0 && all_conns;
    if (all_conns) { // every connection of the frame in ONE round trip: the warps take the limbs in turn, a lane per connection
        for (int l = threadIdx.x >> 5; l < OPP_N_PAIRS; l += blockDim.x >> 5) {
            const int nc = s_nc[l], base = s_coff[l], pa_ofs = s_pofs[s_pa[l]];
            const opp_conn_t *g = p.conns + ((size_t)frame * OPP_N_PAIRS + l) * capP;
            for (int k = threadIdx.x & 31; k < nc; k += 32) {
                const int t = base + k;
                opp_conn_t c;
                c.cid1 = __ldcg(&g[k].cid1), c.cid2 = __ldcg(&g[k].cid2), c.score = __ldcg(&g[k].score);
                s_conn[t] = c;
                if (use_owner) {
                    s_lmb[t] = (unsigned char)l;
                    if (l < 17) {
                        const int la = c.cid1 - pa_ofs;
                        if (la >= 0 && la < capP) s_c1[l * capP + la] = (unsigned short)t;
                        if (c.cid2 >= 0 && c.cid2 < OPP_N_PARTS * capP) atomicOr(&s_in[c.cid2 >> 5], 1u << (c.cid2 & 31));
                    }
                }
            }
        }
    }
    __syncthreads();

    stamp(p, frame, 18, 6);
    const int lane = threadIdx.x & 31;
    int n = 0, hist_max = 0, flags = 0, merges = 0;
    auto peak_score = [&](int id) -> float {
        if (id < 0 || id >= n_peaks) {
            flags |= OPP_FLAG_UB_PEAK_INDEX;
            return 0.f;
        }
        return pk_smem ? __int_as_float(s_pk[id].y) : __ldcg(&peaks[id].score);
    };
    // one limb's connections, in acceptance order; warp 0 only
    auto do_limb = [&](int pair_id, const Span<opp_conn_t> cl, int nconn, int kstart) {
        const int part1 = c_pair_a[pair_id], part2 = c_pair_b[pair_id];
        for (int k = kstart; k < nconn; ++k) {
            const opp_conn_t conn = cl[k];
            // for (auto hr : human_refs) if (hr.touches(...)) hr_ids.push_back(hr.id)   src/paf.cpp:195-199
            int n_hits = 0, hit0 = -1, hit1 = -1;
            for (int base = 0; base < n && n_hits < 2; base += 32) {
                const int q = base + lane;
                const bool t = q < n && (hr[q * HR_WORDS + HR_PART + part1] == conn.cid1 || hr[q * HR_WORDS + HR_PART + part2] == conn.cid2);
                unsigned m = __ballot_sync(0xffffffffu, t);
                while (m && n_hits < 2) {
                    const int b = __ffs(m) - 1;
                    // src/paf.cpp:198 pushes the STORED id (stale after an erase); the Python path's pafprocess
                    // keeps the position in the vector instead
                    const int id = p.true_index ? base + b : hr[(base + b) * HR_WORDS + HR_ID];
                    if (n_hits == 0) hit0 = id; else hit1 = id;
                    ++n_hits;
                    m &= m - 1;
                }
            }
            if (n_hits == 1) {
                if (hit0 < 0 || hit0 >= hist_max) {
                    flags |= OPP_FLAG_UB_STALE_INDEX;
                } else if (lane == 0) {
                    const Span<int> h1 = hr.from(hit0 * HR_WORDS);
                    if (h1[HR_PART + part2] != conn.cid2) {
                        h1[HR_PART + part2] = conn.cid2;
                        h1[HR_NPARTS] += 1;
                        const float sc = __int_as_float(h1[HR_SCORE]);
                        h1[HR_SCORE] = __float_as_int(__fadd_rn(sc, __fadd_rn(peak_score(conn.cid2), conn.score)));
                    }
                }
                flags |= __shfl_sync(0xffffffffu, flags, 0); // keep the UB flag warp-uniform
            } else if (n_hits >= 2) {
                if (hit0 < 0 || hit0 >= hist_max || hit1 < 0 || hit1 >= hist_max) {
                    flags |= OPP_FLAG_UB_STALE_INDEX;
                } else {
                    const Span<int> h1 = hr.from(hit0 * HR_WORDS), h2 = hr.from(hit1 * HR_WORDS);
                    const bool both = lane < OPP_N_PARTS && h1[HR_PART + lane] > 0 && h2[HR_PART + lane] > 0;
                    const bool membership = __ballot_sync(0xffffffffu, both) != 0;
                    if (!membership) {
                        if (lane < OPP_N_PARTS) h1[HR_PART + lane] += h2[HR_PART + lane] + 1;
                        if (lane == 0) {
                            h1[HR_NPARTS] += h2[HR_NPARTS];
                            float sc = __fadd_rn(__int_as_float(h1[HR_SCORE]), __int_as_float(h2[HR_SCORE]));
                            sc = __fadd_rn(sc, conn.score);
                            h1[HR_SCORE] = __float_as_int(sc);
                        }
                        __syncwarp();
                        // human_refs.erase(begin() + hr_ids[1])   src/paf.cpp:231
                        const int e = hit1;
                        if (e >= n) {
                            flags |= OPP_FLAG_UB_ERASE_PAST_END; // libstdc++ 13: nothing moves, size shrinks
                            if (n > 0) --n;
                        } else {
                            // records e+1 .. n-1 move down one record: 128 words per step (a step's loads end before its
                            // stores begin, and a later step only loads words no earlier step has stored)
                            const int w0 = e * HR_WORDS, w1 = (n - 1) * HR_WORDS;
                            for (int wbase = w0; wbase < w1; wbase += 128) {
                                int v[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int wd = wbase + 32 * u + lane;
                                    v[u] = wd < w1 ? hr[wd + HR_WORDS] : 0;
                                }
                                __syncwarp();
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int wd = wbase + 32 * u + lane;
                                    if (wd < w1) hr[wd] = v[u];
                                }
                                __syncwarp();
                            }
                            --n;
                        }
                        ++merges;
                    } else if (lane == 0) {
                        h1[HR_PART + part2] = conn.cid2;
                        h1[HR_NPARTS] += 1;
                        const float sc = __int_as_float(h1[HR_SCORE]);
                        h1[HR_SCORE] = __float_as_int(__fadd_rn(sc, __fadd_rn(peak_score(conn.cid2), conn.score)));
                    }
                }
                flags |= __shfl_sync(0xffffffffu, flags, 0);
            } else if (pair_id <= 16) { // !is_virtual_pair, include/openpose-plus/coco.h:55
                if (n >= capH) {
                    flags |= OPP_FLAG_HUMAN_OVERFLOW;
                } else {
                    const Span<int> hn = hr.from(n * HR_WORDS);
                    if (lane < OPP_N_PARTS) hn[HR_PART + lane] = lane == part1 ? conn.cid1 : (lane == part2 ? conn.cid2 : -1);
                    if (lane == 0) {
                        hn[HR_ID] = n;
                        hn[HR_NPARTS] = 2;
                        const float sc = __fadd_rn(__fadd_rn(peak_score(conn.cid1), peak_score(conn.cid2)), conn.score);
                        hn[HR_SCORE] = __float_as_int(sc);
                    }
                    flags |= __shfl_sync(0xffffffffu, flags, 0);
                    ++n;
                    if (n > hist_max) hist_max = n;
                }
            }
            __syncwarp();
        }
    };

    // The 17 tree limbs (pair_id <= 16) as a forest, whole CTA.  Each of them brings a part that no human holds yet
    // (its second part appears as a second part nowhere earlier and as a first part only later,
    // include/openpose-plus/coco.h:33-53) and the greedy matching uses every peak at most once per limb, so a
    // connection touches at most ONE human (the one holding cid1), a human takes at most one connection per limb, and
    // nothing is erased (stored id == position).  The sequential loop of src/paf.cpp:192-248 therefore reduces to:
    //   * a connection CREATES a human iff nobody holds its cid1 when its limb is reached: no connection of the parent
    //     limb ends in cid1 (the neck, part 1, has no parent limb) and no earlier limb with the same first part (neck:
    //     limbs 0, 1, 6, 9, 12; nose: limbs 13, 15) has a connection from the same peak;
    //   * humans are numbered in (limb, connection) order of their creating connections (ordered compaction);
    //   * a human then collects, limb after limb, the connection that starts at the peak it holds in that limb's first
    //     part -- independent of every other human, so one thread per human walks the limbs with the reference's
    //     score association ((s(cid2) + conn.score) added in limb order).
    auto tree_limbs_forest = [&]() {
        const int T = s_coff[17]; // connections of limbs 0..16, flattened in limb order
        int created = 0;          // humans created so far (uniform over the CTA)
        const Span<int> s_create = s_keep;   // [capH] creating connection of each human (s_keep is not in use yet)
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        int tflags = 0;
        auto pscore = [&](int id) -> float {
            if (id < 0 || id >= n_peaks) {
                tflags |= OPP_FLAG_UB_PEAK_INDEX;
                return 0.f;
            }
            return pk_smem ? __int_as_float(s_pk[id].y) : __ldcg(&peaks[id].score);
        };
        for (int base = 0; base < T; base += blockDim.x) {
            const int t = base + threadIdx.x;
            bool creator = false;
            int l = 0;
            opp_conn_t c;
            c.cid1 = c.cid2 = -1, c.score = 0.f;
            if (t < T) {
                l = s_lmb[t], c = s_conn[t];
                const int pa = s_pa[l], la = c.cid1 - s_pofs[pa];
                bool owned = pa != 1 && c.cid1 >= 0 && c.cid1 < OPP_N_PARTS * capP && ((s_in[c.cid1 >> 5] >> (c.cid1 & 31)) & 1u);
                if (la >= 0 && la < capP) {
                    if (pa == 1) { // earlier limbs rooted at the neck
                        if (l > 0) owned |= s_c1[0 * capP + la] != 0xffff;
                        if (l > 1) owned |= s_c1[1 * capP + la] != 0xffff;
                        if (l > 6) owned |= s_c1[6 * capP + la] != 0xffff;
                        if (l > 9) owned |= s_c1[9 * capP + la] != 0xffff;
                    } else if (l == 15) { // the nose roots limbs 13 and 15: a nose without a neck may already head limb 13's human
                        owned |= s_c1[13 * capP + la] != 0xffff;
                    }
                }
                creator = !owned;
            }
            const unsigned m = __ballot_sync(0xffffffffu, creator);
            if (lane == 0) s_nwarp[warp] = __popc(m);
            __syncthreads();
            int before = created, total = 0;
            for (int q = 0; q < nwarps; ++q) {
                const int cq = s_nwarp[q];
                if (q < warp) before += cq;
                total += cq;
            }
            if (creator) {
                const int pos = before + __popc(m & ((1u << lane) - 1));
                if (pos < capH) { // part ids were preset to -1 by the whole CTA
                    s_create[pos] = t;
                    const Span<int> hq = hr.from(pos * HR_WORDS);
                    hq[HR_ID] = pos;
                    hq[HR_PART + s_pa[l]] = c.cid1, hq[HR_PART + s_pb[l]] = c.cid2;
                }
            }
            // What this connection adds to its human's score, in place of its own score (nothing else reads a tree
            // limb's score from shared memory): the creating one opens the sum with (s(cid1) + s(cid2)) + score
            // (src/paf.cpp:243-244), every other one is added as (s(cid2) + score) (:205-209).
            if (t < T) {
                const float s2 = pscore(c.cid2);
                s_conn[t].score = creator ? __fadd_rn(__fadd_rn(pscore(c.cid1), s2), c.score) : __fadd_rn(s2, c.score);
            }
            created += total;
            __syncthreads();
        }
        if (created > capH) tflags |= OPP_FLAG_HUMAN_OVERFLOW, created = capH;
        // Parts: a human's skeleton below its creating connection is at most three limbs deep (neck - shoulder - elbow -
        // wrist, neck - hip - knee - ankle, neck - nose - eye - ear), so three rounds over (human, limb) settle every part.
        // Limbs before the creating one find nothing: their first part is either not held or heads no connection there.
        for (int round = 0; round < 3; ++round) {
            for (int it = threadIdx.x; it < created * 17; it += blockDim.x) {
                const int q = it / 17, l = it - q * 17;
                const Span<int> hq = hr.from(q * HR_WORDS + HR_PART);
                const int pa = s_pa[l], pb = s_pb[l];
                const int held = hq[pa];
                if (held < 0 || hq[pb] != -1) continue;
                const int la = held - s_pofs[pa];
                if (la < 0 || la >= capP) continue;
                const int t = s_c1[l * capP + la];
                if (t != 0xffff) hq[pb] = s_conn[t].cid2;
            }
            __syncthreads();
        }
        // Score and part count: the creating connection first (src/paf.cpp:243-244), then the human's connection of every
        // later limb in limb order, each as (s(cid2) + conn.score) (:205-209).  The lookups of different limbs are
        // independent now; only the additions are a chain.
        for (int q = threadIdx.x; q < created; q += blockDim.x) {
            const Span<int> hq = hr.from(q * HR_WORDS);
            const int t0 = s_create[q], l0 = s_lmb[t0];
            float sc = s_conn[t0].score;
            int np = 2;
            // no branches: the 17 look-ups are independent of each other and can all be in flight; only the additions chain
#pragma unroll
            for (int l = 0; l < 17; ++l) {
                const int held = hq[HR_PART + c_pair_a[l]];
                const int la = held - s_pofs[c_pair_a[l]];
                const bool in = l != l0 && held >= 0 && la >= 0 && la < capP;
                const int t = s_c1[l * capP + (in ? la : 0)];
                const bool on = in && t != 0xffff;
                const float v = s_conn[on ? t : t0].score;
                sc = on ? __fadd_rn(sc, v) : sc;
                np += on;
            }
            hq[HR_NPARTS] = np, hq[HR_SCORE] = __float_as_int(sc);
        }
        // the flag BITS of all threads (__syncthreads_or only answers "any non-zero")
        if (tflags) atomicOr(&s_state[5], tflags);
        __syncthreads();
        flags |= s_state[5];
        n = created, hist_max = created;
    };

    // A virtual limb (17, 18) while nothing has been erased yet (stored id == position): 32 connections at a time,
    // each lane finds the humans its connection touches.  A connection that touches nobody, or one human that already
    // holds cid2, changes nothing; one human with that part still empty is extended -- and none of these can alter
    // what another connection of the same limb touches (peaks are used once per limb).  Anything else (two humans:
    // merge or overwrite; one human holding another peak there) is order dependent: from the first such connection
    // on, the limb continues in the sequential form.  Returns the index to continue from (nconn: all done).
    auto virtual_limb_prefix = [&](int pair_id, const Span<opp_conn_t> cl, int nconn) -> int {
        const int part1 = c_pair_a[pair_id], part2 = c_pair_b[pair_id];
        for (int base = 0; base < nconn; base += 32) {
            const int k = base + lane;
            const bool on = k < nconn;
            opp_conn_t conn;
            conn.cid1 = conn.cid2 = -2, conn.score = 0.f;
            if (on) conn = cl[k];
            int nh = 0, h0 = 0;
            if (on)
                for (int q = 0; q < n; ++q) {
                    const Span<int> hq = hr.from(q * HR_WORDS + HR_PART);
                    if (hq[part1] == conn.cid1 || hq[part2] == conn.cid2) {
                        if (nh == 0) h0 = q;
                        ++nh;
                    }
                }
            const int cur = (on && nh == 1) ? hr[h0 * HR_WORDS + HR_PART + part2] : -2;
            const bool noop = !on || nh == 0 || (nh == 1 && cur == conn.cid2);
            const bool safe = on && nh == 1 && cur == -1;
            const unsigned mh = __ballot_sync(0xffffffffu, !noop && !safe);
            const int lim = mh ? __ffs(mh) - 1 : 32;
            if (safe && lane < lim) {
                const Span<int> h1 = hr.from(h0 * HR_WORDS);
                h1[HR_PART + part2] = conn.cid2;
                h1[HR_NPARTS] += 1;
                const float sc = __int_as_float(h1[HR_SCORE]);
                h1[HR_SCORE] = __float_as_int(__fadd_rn(sc, __fadd_rn(peak_score(conn.cid2), conn.score)));
            }
            __syncwarp();
            if (mh) return base + lim;
        }
        return nconn;
    };

    if (all_conns) {
        if (use_owner) {
            tree_limbs_forest();
            stamp(p, frame, 18, 10);
        }
        if (threadIdx.x < 32)
            for (int pair_id = use_owner ? 17 : 0; pair_id < OPP_N_PAIRS; ++pair_id) {
                int k0 = 0;
                if (pair_id >= 17 && merges == 0) {
                    k0 = virtual_limb_prefix(pair_id, s_conn.from(s_coff[pair_id]), s_nc[pair_id]);
                    for (int o = 16; o > 0; o >>= 1) flags |= __shfl_xor_sync(0xffffffffu, flags, o); // lanes set UB flags on their own
                }
                do_limb(pair_id, s_conn.from(s_coff[pair_id]), s_nc[pair_id], k0);
            }
    } else { // capacities too large to stage every limb at once: one limb at a time
        for (int pair_id = 0; pair_id < OPP_N_PAIRS; ++pair_id) {
            const opp_conn_t *g = p.conns + ((size_t)frame * OPP_N_PAIRS + pair_id) * capP;
            __syncthreads();
            for (int t = threadIdx.x; t < s_nc[pair_id]; t += blockDim.x) {
                opp_conn_t c;
                c.cid1 = __ldcg(&g[t].cid1), c.cid2 = __ldcg(&g[t].cid2), c.score = __ldcg(&g[t].score);
                s_conn[t] = c;
            }
            __syncthreads();
            if (threadIdx.x < 32) do_limb(pair_id, s_conn, s_nc[pair_id], 0);
        }
    }

    stamp(p, frame, 18, 7);
    // src/paf.cpp:253-260 filter (order kept)
    if (threadIdx.x < 32) {
        int n_out = 0;
        for (int base = 0; base < n; base += 32) {
            const int q = base + lane;
            bool keep = false;
            if (q < n) {
                const int np = hr[q * HR_WORDS + HR_NPARTS];
                const float sc = __int_as_float(hr[q * HR_WORDS + HR_SCORE]);
                keep = !(np < 4 || __fdiv_rn(sc, (float)np) < p.thr_human);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) s_keep[n_out + __popc(m & ((1u << lane) - 1))] = q;
            n_out += __popc(m);
        }
        if (lane == 0) s_state[0] = n, s_state[2] = flags, s_state[3] = merges, s_state[4] = n_out;
    }
    __syncthreads();
    stamp(p, frame, 18, 8);
    // src/paf.cpp:292-310 output: one thread per (human, part)
    const int n_out = s_state[4];
    opp_human_t *out = p.humans + (size_t)frame * capH;
    int *out_parts = p.href_parts + (size_t)frame * capH * OPP_N_PARTS;
    int uflags = 0;
    for (int it = threadIdx.x; it < n_out * OPP_N_PARTS; it += blockDim.x) {
        const int o = it / OPP_N_PARTS, i = it - o * OPP_N_PARTS;
        const int q = s_keep[o];
        const int id = hr[q * HR_WORDS + HR_PART + i];
        out_parts[o * OPP_N_PARTS + i] = id;
        opp_body_part_t bp;
        bp.has_value = 0, bp.pad_[0] = bp.pad_[1] = bp.pad_[2] = 0, bp.x = bp.y = bp.score = 0.f;
        if (id != -1) {
            bp.has_value = 1;
            if (id < 0 || id >= n_peaks) {
                uflags |= OPP_FLAG_UB_PEAK_INDEX;
            } else if (pk_smem) {
                const int2 pk = s_pk[id];
                bp.x = (float)(pk.x & 0xffff), bp.y = (float)(pk.x >> 16), bp.score = __int_as_float(pk.y);
            } else {
                bp.x = (float)__ldcg(&peaks[id].x), bp.y = (float)__ldcg(&peaks[id].y), bp.score = __ldcg(&peaks[id].score);
            }
        }
        out[o].parts[i] = bp;
        if (i == 0) out[o].score = __int_as_float(hr[q * HR_WORDS + HR_SCORE]);
    }
    if (uflags) atomicOr(&s_state[6], uflags);
    __syncthreads();
    if (threadIdx.x == 0) {
        p.n_humans[frame] = n_out;
        const int fl = s_state[2] | s_state[6];
        const int all = atomicOr(p.flags + frame, fl) | fl; // every other writer of this word finished before this CTA became last
        if (p.flags_out) p.flags_out[frame] = all;
        p.stats[frame * 4 + 0] = s_state[0];
        p.stats[frame * 4 + 1] = s_state[3];
        if (p.host_done) {
            // The barrier above orders every thread's records before this point; the fence is cumulative, so they
            // are visible system-wide before the completion word is (barrier, one thread fences, one thread signals).
            __threadfence_system();
            if (atomicAdd(p.batch_done, 1) == p.done_frames - 1) {
                if (p.done_frames > 1) __threadfence_system(); // the other frames' records, observed through the counter
                *reinterpret_cast<volatile int *>(p.host_done) = p.done_tag;
            }
        }
    }
}

// __grid_constant__: the parameter block is read through references (assemble_frame, the stamps); without the qualifier
// taking its address makes every CTA copy it to local memory first
// UNORDERED = p.cand_unordered as a compile-time constant: the default capacities never reach the ordered-compaction
// form of the scoring loop, and the kernel's code is fetched cold (see score_pair) - what is not there is not fetched.
template <bool UNORDERED>
__global__ void __launch_bounds__(OPP_THREADS) k3_limbs(const __grid_constant__ K3Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pair_id = blockIdx.x, frame = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int h = p.g.h, w = p.g.w, capP = p.capP, capC = p.capC;
    const int pa = c_pair_a[pair_id], pb = c_pair_b[pair_id], cx = c_net_x[pair_id]; // the y channel is the next plane
    float *s_paf = reinterpret_cast<float *>(smem_raw + p.off_paf);      // [2][h*w]
    __shared__ unsigned long long s_mbar;
    bool paf_bulk = false;
    // the x and y channels of a limb are adjacent planes: one contiguous tile of 2*h*w floats
    const float *gpx = p.paf + ((size_t)frame * OPP_N_PAF + cx) * h * w;
    auto fetch_paf_tile = [&]() {
        paf_bulk = ((reinterpret_cast<uintptr_t>(gpx) | (uintptr_t)(2 * h * w * sizeof(float))) & 15) == 0;
        if (paf_bulk) {
            if (tid == 0) {
                bulk_init(&s_mbar);
                bulk_load(s_paf, gpx, (unsigned)(2 * h * w * sizeof(float)), &s_mbar);
            }
        } else {
            stage_async(s_paf, gpx, 2 * h * w);
        }
    };
    // Latency path: the PAFs are still in the caller's pinned host memory and nothing the preceding kernels produce is
    // needed to fetch them, so the tile starts its trip over PCIe before this CTA waits for the peak kernel.
    const bool paf_early = p.paf_early && p.paf_in_smem;
    if (paf_early) fetch_paf_tile();
    pdl_wait(); // peaks come from the peak kernel, which may still be running when this CTA is scheduled
    // The two key lists of this limb are fetched whole (capP entries each; those beyond the list's size are ignored)
    // in the same round trip as the list sizes, instead of after them.
    const Span<int> s_ka = SPAN(int, reinterpret_cast<int *>(smem_raw + p.off_keys), 2 * capP), s_kb = s_ka.from(capP);
    for (int t = tid; t < 2 * capP; t += blockDim.x) {
        const int part = t < capP ? pa : pb, k = t < capP ? t : t - capP;
        s_ka[t] = __ldcg(p.pk_key + ((size_t)frame * OPP_N_PARTS + part) * capP + k);
    }
    // sizes of the 18 key lists -> part offsets in all_peaks (one warp scan)
    __shared__ int s_pcnt[OPP_N_PARTS], s_pofs3[OPP_N_PARTS + 1];
    if (tid < 32) {
        const int raw = tid < OPP_N_PARTS ? __ldcg(p.cnt.pk_cnt + frame * OPP_N_PARTS + tid) : 0;
        const int cnt = min(raw, capP);
        if (raw > capP && (tid == pa || tid == pb)) atomicOr(p.flags + frame, OPP_FLAG_PEAK_OVERFLOW);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        if (tid < OPP_N_PARTS) s_pcnt[tid] = cnt, s_pofs3[tid + 1] = incl;
        if (tid == 0) s_pofs3[0] = 0;
    }
    __syncthreads();
    const int ofs_a = s_pofs3[pa], na = s_pcnt[pa];
    const int ofs_b = s_pofs3[pb], nb = s_pcnt[pb];
    opp_peak_t *peaks = p.peaks + (size_t)frame * OPP_N_PARTS * capP;

    const Span<int2> s_pa = SPAN(int2, reinterpret_cast<int2 *>(smem_raw + p.off_pk), capP);          // [capP]
    const Span<int2> s_pb = SPAN(int2, reinterpret_cast<int2 *>(smem_raw + p.off_pk) + capP, capP);   // [capP]
    const Span<unsigned char> s_used = SPAN(unsigned char, smem_raw + p.off_used, 2 * capP);         // [2*capP]
    const Span<int> s_misc = SPAN(int, reinterpret_cast<int *>(smem_raw + p.off_misc), 16);          // [16]: warp counts, totals
    Cand *cand0_, *cand1_;
    if (p.cand_in_smem) {
        cand0_ = reinterpret_cast<Cand *>(smem_raw + p.off_cand);
        cand1_ = reinterpret_cast<Cand *>(smem_raw + p.off_cand1); // the PAF tile's bytes: written only after the scoring loop
    } else {
        cand0_ = reinterpret_cast<Cand *>(p.cand_scratch) + ((size_t)frame * OPP_N_PAIRS + pair_id) * 2 * capC;
        cand1_ = cand0_ + capC;
    }
    const Span<Cand> cand0 = SPAN(Cand, cand0_, capC), cand1 = SPAN(Cand, cand1_, capC);

    stamp(p, frame, pair_id, 0);
    if (p.times && tid == 0) { // which SM ran this CTA (slot 11)
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        p.times[((size_t)frame * OPP_N_PAIRS + pair_id) * 12 + 11] = sm + 1;
    }
    // this limb's two parts: keys -> raster order (all_peaks slices in global memory, (x, y) lists in shared memory)
    {
        PeakSource src;
        src.g = p.g, src.conf = p.conf, src.conf_up = p.conf_up;
        order_part_peaks(src, frame, pa, s_ka, na, ofs_a, s_pa, c_part_writer[pa] == pair_id ? peaks : nullptr);
        order_part_peaks(src, frame, pb, s_kb, nb, ofs_b, s_pb, c_part_writer[pb] == pair_id ? peaks : nullptr);
    }
    int n_cand = 0;
    const long n_pairs = (long)na * nb;
    if (n_pairs > 0) {
        const float *gpy = gpx + h * w;
        const float *px_plane = gpx, *py_plane = gpy;
        if (p.paf_in_smem) {
            if (!paf_early) fetch_paf_tile();
            px_plane = s_paf, py_plane = s_paf + h * w;
        }
        for (int t = tid; t < 2 * capP; t += blockDim.x) s_used[t] = 0;
        // the step table and the weak-cell map pay for themselves from a few hundred pairs on (a frame with five people has
        // ~25 pairs per limb: there they would only add to the one-frame latency)
        const bool many_pairs = n_pairs >= 512;
        if (p.steps_in_smem && many_pairs) { // d / 10.f for every coordinate difference the image allows
            float *st = reinterpret_cast<float *>(smem_raw + p.off_steps);
            for (int t = tid; t < max(p.g.H, p.g.W); t += blockDim.x) st[t] = __fdiv_rn((float)t, 10.f);
        }
        if (tid < 16) s_misc[tid] = 0;
        stage_wait();
        __syncthreads(); // also publishes the mbarrier initialisation to the waiting threads
        if (paf_bulk) bulk_wait(&s_mbar, 0);
        stamp(p, frame, pair_id, 1);

        PairCtx pc;
        pc.px = SPAN(const float, px_plane, h * w), pc.py = SPAN(const float, py_plane, h * w), pc.g = &p.g, pc.w = w, pc.H = p.g.H, pc.thr = p.thr_vec;
        pc.sshift = (p.g.S > 0 && (p.g.S & (p.g.S - 1)) == 0) ? 31 - __clz(p.g.S) : -1;
        pc.steps = SPAN(const float, p.steps_in_smem && many_pairs ? reinterpret_cast<const float *>(smem_raw + p.off_steps) : nullptr, max(p.g.H, p.g.W));
        pc.weak = SPAN(const unsigned, nullptr, 0);
        if (p.weak_in_smem && many_pairs && pc.sshift >= 0 && p.g.W < (1 << 22) && p.g.H < (1 << 22)) {
            // cells whose PAF is too small for any sample to pass (see pair_may_pass)
            const Span<unsigned> wk = SPAN(unsigned, reinterpret_cast<unsigned *>(smem_raw + p.off_weak), (h * w + 31) / 32);
            const float tw = __fmul_rn(p.thr_vec, 1.f - 1.f / 8192.f);
            for (int base = warp * 32; base < h * w; base += blockDim.x) {
                const int cidx = base + lane;
                const bool wkb = cidx >= h * w || __fadd_rn(fabsf(px_plane[cidx]), fabsf(py_plane[cidx])) <= tw;
                const unsigned m = __ballot_sync(0xffffffffu, wkb);
                if (lane == 0) wk[base >> 5] = m;
            }
            pc.weak = SPAN(const unsigned, wk.p, (h * w + 31) / 32);
            __syncthreads();
        }
        const Span<int> s_surv = SPAN(int, reinterpret_cast<int *>(smem_raw + p.off_surv), p.surv_cap); // [surv_cap] pairs that passed the quick test
        const unsigned n_pairs_u = (unsigned)n_pairs, nb_u = (unsigned)nb;
        const unsigned lt = (1u << lane) - 1;
        // idx -> (ia, ib) without an integer division: q = floor(idx * floor(2^32 / nb) / 2^32) is ia or ia - 1
        const unsigned nb_inv = nb_u > 1 ? (unsigned)(0x100000000ull / nb_u) : 0u;
        auto split = [&](unsigned idx, unsigned &ia, unsigned &ib) {
            if (nb_u == 1) {
                ia = idx, ib = 0;
                return;
            }
            ia = __umulhi(idx, nb_inv);
            ib = idx - ia * nb_u;
            if (ib >= nb_u) ib -= nb_u, ++ia;
        };
        int overflow = 0, n_surv_total = 0;
        if (UNORDERED) {
            // Candidates are appended in whatever order the warps finish (one shared-memory atomic per warp and round,
            // no block-wide compaction); the sort below ranks them by (score, pair index), which is std::sort's
            // result whenever no two scores are equal, and restores the a-major / b-minor input order first when some are.
            const Span<int> s_cnt = s_misc.from(12); // [0] survivors in the list, [1] candidates appended
            if (tid == 0) s_cnt[0] = 0, s_cnt[1] = 0;
            __syncthreads();
            // direct = true: no list, the first n_direct pairs themselves (limbs with few pairs skip the filter stage)
            auto drain = [&](bool direct, int n_direct) { // the listed survivors in full; whole CTA
                __syncthreads();
                const int n_surv = direct ? n_direct : min(s_cnt[0], p.surv_cap);
                n_surv_total += n_surv;
                for (int base = warp * 32; base < n_surv; base += blockDim.x) {
                    const int t = base + lane;
                    bool accept = false;
                    float crit2 = 0.f;
                    unsigned ia = 0, ib = 0;
                    if (t < n_surv) {
                        split(direct ? (unsigned)t : (unsigned)s_surv[t], ia, ib);
                        accept = score_pair<false>(pc, s_pa[ia], s_pb[ib], crit2);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, accept);
                    if (m) {
                        int base_pos = 0;
                        if (lane == 0) base_pos = atomicAdd(&s_cnt[1], __popc(m));
                        base_pos = __shfl_sync(0xffffffffu, base_pos, 0);
                        if (accept) {
                            const int pos = base_pos + __popc(m & lt);
                            if (pos < capC) {
                                Cand cd;
                                cd.i1 = ofs_a + (int)ia, cd.i2 = ofs_b + (int)ib, cd.s = crit2;
                                cand0[pos] = cd;
                            } else
                                overflow = 1;
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) s_cnt[0] = 0;
                __syncthreads();
            };
            const unsigned chunk = (unsigned)(p.surv_cap / 2) & ~255u; // pairs per round of quick tests: the list never overflows
            unsigned listed_max = 0;                                    // upper bound of the survivors listed so far
            for (unsigned c0 = 0; many_pairs && c0 < n_pairs_u; c0 += chunk) {
                const unsigned c1 = min(c0 + chunk, n_pairs_u);
                if (listed_max + (c1 - c0) > (unsigned)p.surv_cap) {
                    drain(false, 0);
                    listed_max = 0;
                }
                for (unsigned base = c0 + warp * 32; base < c1; base += blockDim.x) {
                    const unsigned idx = base + lane;
                    bool alive = false;
                    if (idx < c1) {
                        unsigned ia, ib;
                        split(idx, ia, ib);
                        float unused;
                        alive = pc.weak.p ? pair_may_pass(pc, s_pa[ia], s_pb[ib]) : score_pair<true>(pc, s_pa[ia], s_pb[ib], unused);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, alive);
                    if (m) {
                        int base_pos = 0;
                        if (lane == 0) base_pos = atomicAdd(&s_cnt[0], __popc(m));
                        base_pos = __shfl_sync(0xffffffffu, base_pos, 0);
                        if (alive) s_surv[base_pos + __popc(m & lt)] = (int)idx;
                    }
                }
                listed_max += c1 - c0;
            }
            drain(!many_pairs, (int)n_pairs_u);
            n_cand = s_cnt[1];
        } else {
        // ordered block compaction: the threads with `flag` learn their position after the `count` entries already
        // there, in thread order (candidates must stay a-major / b-minor: it is std::sort's input order)
        auto place = [&](bool flag, int count, int &pos) -> int {
            const unsigned m = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) s_misc[warp] = __popc(m);
            __syncthreads();
            int before = count, total = 0;
            for (int q = 0; q < nwarps; ++q) {
                const int cq = s_misc[q];
                if (q < warp) before += cq;
                total += cq;
            }
            pos = before + __popc(m & lt);
            return total;
        };
        unsigned next = 0;
        while (next < n_pairs_u) {
            // (1) quick test of the next pairs until the survivor list is (nearly) full
            int n_surv = 0;
            while (next < n_pairs_u && n_surv + (int)blockDim.x <= p.surv_cap) {
                const unsigned idx = next + tid;
                bool alive = false;
                if (idx < n_pairs_u) {
                    unsigned ia, ib;
                    split(idx, ia, ib);
                    float unused;
                    alive = score_pair<true>(pc, s_pa[ia], s_pb[ib], unused);
                }
                int pos;
                const int total = place(alive, n_surv, pos);
                if (alive) s_surv[pos] = (int)idx;
                n_surv += total;
                next += blockDim.x;
                __syncthreads();
            }
            n_surv_total += n_surv;
            // (2) the survivors in full
            for (int base = 0; base < n_surv; base += blockDim.x) {
                const int t = base + tid;
                bool accept = false;
                float crit2 = 0.f;
                unsigned ia = 0, ib = 0;
                if (t < n_surv) {
                    split((unsigned)s_surv[t], ia, ib);
                    accept = score_pair<false>(pc, s_pa[ia], s_pb[ib], crit2);
                }
                int pos;
                const int total = place(accept, n_cand, pos);
                if (accept) {
                    if (pos < capC) {
                        Cand cd;
                        cd.i1 = ofs_a + (int)ia, cd.i2 = ofs_b + (int)ib, cd.s = crit2;
                        cand0[pos] = cd;
                    } else
                        overflow = 1;
                }
                n_cand += total;
                __syncthreads();
            }
        }
        }
        if (tid == 0) atomicAdd(p.stats + frame * 4 + 3, n_surv_total);
        if (__syncthreads_or(overflow)) {
            if (tid == 0) atomicOr(p.flags + frame, OPP_FLAG_CAND_OVERFLOW);
        }
        n_cand = min(n_cand, capC);
        stamp(p, frame, pair_id, 2);

        // ---- sort by score, descending, in std::sort's order.  With no equal scores the sorted order is
        // unique and a parallel rank sort gives it; with ties only the sequential emulation does.
        const Span<Cand> sorted = sort_candidates_desc(cand0, cand1, n_cand, UNORDERED, ofs_a, ofs_b, nb_u, p.cand_in_smem != 0,
                                                       SPAN(int, reinterpret_cast<int *>(smem_raw + p.off_surv), p.surv_cap));

        stamp(p, frame, pair_id, 3);
        // ---- greedy matching in sorted order (src/paf.cpp:154-173): a candidate is accepted unless an accepted one
        // before it shares a peak.  Parallel form: among the candidates still undecided, one that is the FIRST (in
        // sorted order) for both of its peaks has no undecided predecessor sharing a peak - and none of the accepted
        // ones does, or it would have been rejected - so the sequential loop accepts it; accepting it rejects every
        // other undecided candidate on its two peaks.  Rounds of (first-per-peak by atomicMin, accept, reject) decide
        // all candidates in a handful of rounds on real frames; the position in sorted order is the only priority, so
        // ties in the scores change nothing here.  Whatever is still undecided after GREEDY_ROUNDS is finished by the
        // sequential loop; the accepted candidates are then written in sorted (= acceptance) order.
        const Span<opp_conn_t> conns = SPAN(opp_conn_t, p.conns + ((size_t)frame * OPP_N_PAIRS + pair_id) * capP, capP);
        constexpr int GREEDY_ROUNDS = 6;
        const Span<unsigned char> s_state = SPAN(unsigned char, reinterpret_cast<unsigned char *>(smem_raw + p.off_surv), p.surv_cap * 4); // [n_cand] 0 undecided, 1 accepted, 2 rejected
        if (n_cand > 32 && n_cand <= p.surv_cap * 4) {
            const Span<int> s_first = s_ka; // [2 capP] first undecided candidate of every peak (the key lists are dead by now)
            for (int t = tid; t < n_cand; t += blockDim.x) s_state[t] = 0;
            int left = 1;
            for (int round = 0; round < GREEDY_ROUNDS && left; ++round) {
                for (int t = tid; t < 2 * capP; t += blockDim.x) s_first[t] = 0x7fffffff;
                __syncthreads();
                for (int t = tid; t < n_cand; t += blockDim.x)
                    if (!s_state[t]) {
                        atomicMin(&s_first[sorted[t].i1 - ofs_a], t);
                        atomicMin(&s_first[capP + sorted[t].i2 - ofs_b], t);
                    }
                __syncthreads();
                for (int t = tid; t < n_cand; t += blockDim.x)
                    if (!s_state[t]) {
                        const int la = sorted[t].i1 - ofs_a, lb = sorted[t].i2 - ofs_b;
                        if (s_first[la] == t && s_first[capP + lb] == t) s_state[t] = 1, s_used[la] = 1, s_used[capP + lb] = 1;
                    }
                __syncthreads();
                left = 0;
                for (int t = tid; t < n_cand; t += blockDim.x)
                    if (!s_state[t]) {
                        if (s_used[sorted[t].i1 - ofs_a] | s_used[capP + sorted[t].i2 - ofs_b]) s_state[t] = 2;
                        else left = 1;
                    }
                left = __syncthreads_or(left);
            }
            if (left) { // long chains of candidates, each blocked only by its predecessor: the rest in sequence
                if (tid == 0)
                    for (int t = 0; t < n_cand; ++t)
                        if (!s_state[t]) {
                            const int la = sorted[t].i1 - ofs_a, lb = sorted[t].i2 - ofs_b;
                            if (s_used[la] | s_used[capP + lb]) continue;
                            s_used[la] = 1, s_used[capP + lb] = 1, s_state[t] = 1;
                        }
                __syncthreads();
            }
            int nc = 0;
            for (int base = 0; base < n_cand; base += blockDim.x) {
                const int t = base + tid;
                const bool acc = t < n_cand && s_state[t] == 1;
                const unsigned m = __ballot_sync(0xffffffffu, acc);
                if (lane == 0) s_misc[warp] = __popc(m);
                __syncthreads();
                int before = nc, total = 0;
                for (int q = 0; q < nwarps; ++q) {
                    const int cq = s_misc[q];
                    if (q < warp) before += cq;
                    total += cq;
                }
                if (acc) {
                    const Cand cd = sorted[t];
                    opp_conn_t cn;
                    cn.cid1 = cd.i1, cn.cid2 = cd.i2, cn.score = cd.s;
                    conns[before + __popc(m & ((1u << lane) - 1))] = cn; // at most min(na, nb) <= capP are accepted
                }
                nc += total;
                __syncthreads();
            }
            if (tid == 0) {
                p.n_conns[frame * OPP_N_PAIRS + pair_id] = nc;
                atomicAdd(p.stats + frame * 4 + 2, n_cand);
            }
        } else if (tid == 0) {
            int nc = 0;
            for (int t = 0; t < n_cand; ++t) {
                const Cand cd = sorted[t];
                const int la = cd.i1 - ofs_a, lb = cd.i2 - ofs_b;
                if (s_used[la] | s_used[capP + lb]) continue;
                s_used[la] = 1, s_used[capP + lb] = 1;
                opp_conn_t cn;
                cn.cid1 = cd.i1, cn.cid2 = cd.i2, cn.score = cd.s;
                conns[nc++] = cn;
            }
            p.n_conns[frame * OPP_N_PAIRS + pair_id] = nc;
            atomicAdd(p.stats + frame * 4 + 2, n_cand);
        }
    } else {
        if (tid == 0) p.n_conns[frame * OPP_N_PAIRS + pair_id] = 0;
        if (paf_early) { // the tile was requested anyway: it must have landed before the shared memory is reused
            stage_wait();
            __syncthreads();
            if (paf_bulk) bulk_wait(&s_mbar, 0);
        }
    }

    stamp(p, frame, pair_id, 4);
    if (tile_done_is_last(p.cnt.k3_done + frame, OPP_N_PAIRS)) {
        stamp(p, frame, pair_id, 5);
        assemble_frame(p, frame, smem_raw, s_pofs3);
        stamp(p, frame, pair_id, 9);
    }
}
} // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int g_max_smem = 48 * 1024;
static int g_sm_count = 148;

// opt in to the large dynamic shared-memory carve-out (the per-block limit counts static shared memory too)
template <typename Kern> static cudaError_t allow_big_smem(Kern kern, int *dyn_limit)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    *dyn_limit = g_max_smem - (int)fa.sharedSizeBytes;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, *dyn_limit);
}

// The attribute is per kernel AND per device (a process may hold handles on several GPUs): each
// launcher caches the limit per device ordinal.
#define OPP_MAX_DEVICES 64
#define BIG_SMEM_LIMIT(kern, out)                                                   \
    do {                                                                            \
        static int cache_[OPP_MAX_DEVICES];                                         \
        int dev_ = 0;                                                               \
        if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= OPP_MAX_DEVICES) dev_ = 0; \
        if (!cache_[dev_]) {                                                        \
            cudaError_t e_ = allow_big_smem(kern, &cache_[dev_]);                   \
            if (e_ != cudaSuccess) return e_;                                       \
        }                                                                           \
        out = cache_[dev_];                                                         \
    } while (0)

// Launch with or without the programmatic-stream-serialization attribute (see pdl_wait above).
template <typename P> static cudaError_t launch_ex(void (*kern)(const P), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const P &p)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int S, int R, bool STORE, bool BZ>
static cudaError_t launch_k2_fast_t(const K2Params &p, int n_frames, size_t smem, cudaStream_t st, bool pdl)
{
    constexpr int NB = R > S ? 2 : 1;
    int dyn_limit = 0;
    BIG_SMEM_LIMIT((k2_peaks_fast<S, R, STORE, BZ>), dyn_limit);
    if (smem > (size_t)dyn_limit) return cudaErrorInvalidValue;
    dim3 grid(p.nxs * p.nys, STORE ? OPP_N_HEAT : OPP_N_PARTS, n_frames);
    const int groups = (S * p.tw + 61) / 62;
    if (groups < 1 || groups > K2_FAST_MAX_WARPS || p.th + 2 + 2 * NB > 64 || p.tw + 2 * NB > 64) return cudaErrorInvalidValue; // 64-bit activity masks
    if (p.g.h < NB + 1 || p.g.w < NB + 1) return cudaErrorInvalidValue; // make_win's border operands
    const int threads = 32 * groups;
    if (STORE && 2 * ((S * p.tw) >> 2) > threads) return cudaErrorInvalidValue; // two threads per 16-byte column of the tile
    return launch_ex(k2_peaks_fast<S, R, STORE, BZ>, grid, dim3(threads), smem, st, pdl, p);
}

template <int S, int R, bool BZ>
static cudaError_t launch_k2_fast_sr(const K2Params &p, int n_frames, size_t smem, cudaStream_t st, bool pdl)
{
    return p.up_conf ? launch_k2_fast_t<S, R, true, BZ>(p, n_frames, smem, st, pdl) : launch_k2_fast_t<S, R, false, BZ>(p, n_frames, smem, st, pdl);
}

bool k2_fast_supported(const OppGeom &g, bool border_zero)
{
    // integer scale 8 or 4 on both axes, Gaussian radius within two feature cells (R > S needs maps of 3 x 3 cells);
    // the zero-border form is compiled for the Python graph's own size only (x8, k = 25)
    if (border_zero) return g.S == 8 && g.R == 12 && g.h >= 3 && g.w >= 3;
    const int nb = g.R > g.S ? 2 : 1;
    return (g.S == 8 || g.S == 4) && g.R <= 2 * g.S && g.h >= nb + 1 && g.w >= nb + 1;
}

size_t k2_fast_smem_bytes(const OppGeom &g, int tw, int th)
{
    const int nb = g.R > g.S ? 2 : 1;
    const int nr = (th + 2 + 2 * nb < g.h) ? th + 2 + 2 * nb : g.h;
    const int ncol = (tw + 2 * nb < g.w) ? tw + 2 * nb : g.w;
    size_t fl = ((size_t)nr * g.w + 3) & ~(size_t)3;
    fl += ((size_t)nr * g.S * ncol + 3) & ~(size_t)3;
    fl += 2 * (size_t)(th < g.h ? th : g.h) * g.w; // PAF feature rows of the fused-store variant
    return fl * sizeof(float);
}

cudaError_t launch_k2_fast(const K2Params &p, int n_frames, cudaStream_t st, bool pdl)
{
    const size_t smem = k2_fast_smem_bytes(p.g, p.tw, p.th);
    size_t need = smem;
    if ((size_t)OPP_N_PARTS * p.capP * sizeof(int) > need) need = (size_t)OPP_N_PARTS * p.capP * sizeof(int);
    if (need > (size_t)g_max_smem) return cudaErrorInvalidValue;
    if (p.border_zero) { // the Python graph: 'SAME' zero padding, fixed k = 25 (post_process.py:13-32)
        if (p.g.S == 8 && p.g.R == 12) return launch_k2_fast_sr<8, 12, true>(p, n_frames, need, st, pdl);
        return cudaErrorInvalidValue;
    }
#define K2_CASE(S_, R_)                                                       \
    if (p.g.S == S_ && p.g.R == R_) return launch_k2_fast_sr<S_, R_, false>(p, n_frames, need, st, pdl);
    K2_CASE(8, 8) K2_CASE(8, 6) K2_CASE(8, 4)               // k = 17, 13, 9: the kernel sizes the reference's demos and scripts use
    K2_CASE(8, 7) K2_CASE(8, 5) K2_CASE(8, 3) K2_CASE(8, 2) K2_CASE(8, 1) K2_CASE(8, 0)
    K2_CASE(4, 4) K2_CASE(4, 3) K2_CASE(4, 2) K2_CASE(4, 1) K2_CASE(4, 0)
    K2_CASE(8, 12)                                          // k = 25, the Python graph's size, under cv::GaussianBlur's border
    K2_CASE(8, 9) K2_CASE(8, 10) K2_CASE(8, 11) K2_CASE(8, 13) K2_CASE(8, 14) K2_CASE(8, 15) K2_CASE(8, 16)
    K2_CASE(4, 5) K2_CASE(4, 6) K2_CASE(4, 7) K2_CASE(4, 8)
#undef K2_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_k2_generic(const K2Params &p, int n_frames, cudaStream_t st)
{
    const int R = p.g.R;
    const int IW = G_TX + 2 + 2 * R, IH = G_TY + 2 + 2 * R, TW = G_TX + 2;
    // feature rows under the IH image rows of a region (+ the second tap's row, + rounding): they are resized
    // horizontally into the bytes the row-pass result and the smoothed tile use afterwards
    long nfr = ((long)IH * p.g.h + p.g.H - 1) / p.g.H + 3;
    if (nfr > p.g.h) nfr = p.g.h;
    size_t after_in = (size_t)IH * TW + (size_t)(G_TY + 2) * TW;
    if ((size_t)nfr * IW > after_in) after_in = (size_t)nfr * IW;
    size_t smem = ((size_t)IH * IW + after_in) * sizeof(float);
    if ((size_t)OPP_N_PARTS * p.capP * sizeof(int) > smem) smem = (size_t)OPP_N_PARTS * p.capP * sizeof(int);
    int dyn_limit = 0;
    BIG_SMEM_LIMIT(k2_peaks_generic, dyn_limit);
    if (smem > (size_t)dyn_limit) return cudaErrorInvalidValue;
    const int tiles = ((p.g.W + G_TX - 1) / G_TX) * ((p.g.H + G_TY - 1) / G_TY);
    dim3 grid(tiles, OPP_N_PARTS, n_frames);
    k2_peaks_generic<<<grid, OPP_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_k2_generic_rep(const K2Params &p, int n_frames, cudaStream_t st)
{
    const int R = p.g.R, S = p.g.S;
    if (S < 1) return cudaErrorInvalidValue;
    const int IW = G_TX + 2 + 2 * R, IH = G_TY + 2 + 2 * R, TW = G_TX + 2;
    const int nfr = (IH + S - 1) / S + 1; // most feature rows a region can touch
    size_t smem = ((size_t)nfr * IW + (size_t)(nfr + 1) * TW + (size_t)(G_TY + 2) * TW) * sizeof(float) + (size_t)IH * sizeof(int);
    if ((size_t)OPP_N_PARTS * p.capP * sizeof(int) > smem) smem = (size_t)OPP_N_PARTS * p.capP * sizeof(int);
    int dyn_limit = 0;
    BIG_SMEM_LIMIT(k2_peaks_generic_rep, dyn_limit);
    if (smem > (size_t)dyn_limit) return cudaErrorInvalidValue;
    const int tiles = ((p.g.W + G_TX - 1) / G_TX) * ((p.g.H + G_TY - 1) / G_TY);
    dim3 grid(tiles, OPP_N_PARTS, n_frames);
    k2_peaks_generic_rep<<<grid, OPP_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_k3(const K3Params &p, int n_frames, size_t smem, cudaStream_t st, bool pdl)
{
    int dyn_limit = 0;
    if (p.cand_unordered) BIG_SMEM_LIMIT(k3_limbs<true>, dyn_limit);
    else BIG_SMEM_LIMIT(k3_limbs<false>, dyn_limit);
    if (smem > (size_t)dyn_limit) return cudaErrorInvalidValue;
    dim3 grid(OPP_N_PAIRS, n_frames);
    // Threads per limb CTA (OPP_K3_THREADS = 128 .. 256, multiple of 32, overrides).  Most of the kernel is latency bound and
    // leaves lanes idle (a frame with five people has 25 pairs per limb); 192-thread CTAs measured 4-5 % faster than 256 on
    // typical and crowded batches alike (more CTAs fit beside the peak kernel's, whose registers are what keeps them out).
    constexpr int K3_THREADS_DEFAULT = 192;
    static const int threads = [] {
        const char *e = getenv("OPP_K3_THREADS");
        const int t = e ? atoi(e) : K3_THREADS_DEFAULT;
        return (t >= 128 && t <= OPP_THREADS && t % 32 == 0) ? t : K3_THREADS_DEFAULT;
    }();
    return p.cand_unordered ? launch_ex(k3_limbs<true>, grid, dim3(threads), smem, st, pdl, p) : launch_ex(k3_limbs<false>, grid, dim3(threads), smem, st, pdl, p);
}

static bool k1_fast_ok(const K1Params &p)
{
    const OppGeom &g = p.g;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.dst) | reinterpret_cast<uintptr_t>(p.dst2)) & 15) == 0;
    return p.layout == OPP_LAYOUT_CHW && g.S == 8 && (g.W & 3) == 0 && aligned;
}

cudaError_t launch_k1(const K1Params &p, cudaStream_t st)
{
    const OppGeom &g = p.g;
    if (p.layout == OPP_LAYOUT_HWC && g.S == 8) {
        const int Cmax = p.C2 > p.C ? p.C2 : p.C;
        const size_t smem = ((size_t)g.W * Cmax + (size_t)Cmax * g.w) * sizeof(float);
        const bool ok16 = (((size_t)g.W * p.C * 4) & 15) == 0 && (p.C2 == 0 || (((size_t)g.W * p.C2 * 4) & 15) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.dst) | reinterpret_cast<uintptr_t>(p.dst2)) & 15) == 0;
        int dyn_limit = 0;
        BIG_SMEM_LIMIT(k1_replicate_hwc<8>, dyn_limit);
        if (ok16 && smem <= (size_t)dyn_limit) {
            dim3 grid(g.h, p.C2 ? 2 : 1, p.n);
            k1_replicate_hwc<8><<<grid, OPP_THREADS, smem, st>>>(p);
            return cudaGetLastError();
        }
    }
    static const int mode = getenv("OPP_K1_MODE") ? atoi(getenv("OPP_K1_MODE")) : 2; // 0 rows, 1 TMA bulk, 2 first version
    if (k1_fast_ok(p) && mode == 0 && 2 * (g.W >> 2) <= 224) {
        static const int G = getenv("OPP_K1_G") ? atoi(getenv("OPP_K1_G")) : 4; // feature rows per item: 32 output rows, 55 KB contiguous
        static const int nc = getenv("OPP_K1_CTAS") ? atoi(getenv("OPP_K1_CTAS")) : 3;
        const long n_items = (long)((g.h + G - 1) / G) * (p.C + p.C2) * p.n;
        long grid = (long)g_sm_count * nc;
        if (grid > n_items) grid = n_items;
        k1_replicate_rows<8><<<(unsigned)grid, 224, 0, st>>>(p, G, (int)n_items);
        return cudaGetLastError();
    }
    if (k1_fast_ok(p) && mode == 1) {
        const int G = 8; // feature rows per item: 8 x 8 rows of W floats = 64 bulk copies, 2 x 13.8 KB shared at W = 432
        const int groups = (g.h + G - 1) / G;
        const long n_items = (long)groups * (p.C + p.C2) * p.n;
        const size_t smem = 2 * (size_t)G * g.W * sizeof(float);
        static int ctas_per_sm = getenv("OPP_K1_CTAS") ? atoi(getenv("OPP_K1_CTAS")) : 2;
        long grid = (long)g_sm_count * ctas_per_sm;
        if (grid > n_items) grid = n_items;
        if (smem <= 48 * 1024) {
            k1_replicate_bulk<8><<<(unsigned)grid, 128, smem, st>>>(p, G, (int)n_items);
            return cudaGetLastError();
        }
    }
    if (k1_fast_ok(p)) {
        const int rows = 4; // source rows per CTA -> 32 output rows
        dim3 grid((g.h + rows - 1) / rows, p.C + p.C2, p.n);
        k1_replicate_chw<8><<<grid, OPP_THREADS, rows * g.w * sizeof(float), st>>>(p, rows);
        return cudaGetLastError();
    }
    if (p.layout == OPP_LAYOUT_CHW && g.S <= 0 && g.xofs) { // non-integer scale, channels-first: row-buffer form
        const long nfr_max = std::min<long>(g.h, ((long)K1G_ROWS * g.h + g.H - 1) / g.H + 3);
        const size_t smem = (size_t)nfr_max * g.W * sizeof(float);
        int dyn_limit = 0;
        BIG_SMEM_LIMIT(k1_general_chw, dyn_limit);
        if (smem <= (size_t)dyn_limit) {
            dim3 grid((g.H + K1G_ROWS - 1) / K1G_ROWS, p.C + p.C2, p.n);
            k1_general_chw<<<grid, OPP_THREADS, smem, st>>>(p);
            return cudaGetLastError();
        }
    }
    K1Params q = p;
    for (int part = 0; part < 2; ++part) {
        if (part == 1) {
            if (!p.C2) break;
            q.src = p.src2, q.dst = p.dst2, q.C = p.C2;
        }
        q.src2 = nullptr, q.dst2 = nullptr, q.C2 = 0;
        const size_t total = (size_t)q.n * q.C * g.H * g.W;
        size_t blocks = (total + OPP_THREADS - 1) / OPP_THREADS;
        if (blocks > 148 * 64) blocks = 148 * 64;
        k1_general<<<(unsigned)blocks, OPP_THREADS, 0, st>>>(q);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_hwc_to_chw(const float *src, float *dst, int n, int C, int h, int w, cudaStream_t st)
{
    const size_t total = (size_t)n * C * h * w;
    size_t blocks = (total + OPP_THREADS - 1) / OPP_THREADS;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k0_hwc_to_chw<<<(unsigned)blocks, OPP_THREADS, 0, st>>>(src, dst, n, C, h * w);
    return cudaGetLastError();
}

cudaError_t launch_ingest(const float *src0, float *dst0, size_t n0, const float *src1, float *dst1, size_t n1, int *counters, int n_counters,
                          cudaStream_t st)
{
    size_t blocks = ((n0 + n1) / 4 + OPP_THREADS - 1) / OPP_THREADS;
    if (blocks > (size_t)g_sm_count * 4) blocks = (size_t)g_sm_count * 4;
    if (blocks < 1) blocks = 1;
    k0_ingest<<<(unsigned)blocks, OPP_THREADS, 0, st>>>(src0, dst0, n0, src1, dst1, n1, counters, n_counters);
    return cudaGetLastError();
}

// First bounds violation recorded by a -DOPP_DEBUG_BOUNDS build on the current device: {source line of the Span, index,
// size, number of violations}; cudaErrorNotSupported in a release build.
cudaError_t launch_debug_sort(void *cands, int n, int mode, int threads, cudaStream_t st)
{
    if (n < 0 || (mode == 0 && n > 4096) || threads < 32 || threads > OPP_THREADS || threads % 32) return cudaErrorInvalidValue;
    const size_t smem = mode == 0 ? 2 * (size_t)n * sizeof(Cand) + 4 * ((size_t)n / 17 + 2) * sizeof(int) + 16 : 0;
    int dyn_limit = 0;
    BIG_SMEM_LIMIT(k_debug_sort, dyn_limit);
    if (smem > (size_t)dyn_limit) return cudaErrorInvalidValue;
    k_debug_sort<<<1, threads, smem, st>>>(reinterpret_cast<Cand *>(cands), n, mode);
    return cudaGetLastError();
}

cudaError_t opp_kernels_bounds_report(int out[4], bool reset)
{
#ifdef OPP_DEBUG_BOUNDS
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyFromSymbol(out, g_oob, 4 * sizeof(int));
    if (e != cudaSuccess || !reset) return e;
    const int zero[4] = {0, 0, 0, 0};
    return cudaMemcpyToSymbol(g_oob, zero, sizeof zero);
#else
    (void)out, (void)reset;
    return cudaErrorNotSupported;
#endif
}

cudaError_t opp_kernels_init(int max_smem_optin)
{
    g_max_smem = max_smem_optin;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_sm_count = sms;
    return cudaSuccess;
}
