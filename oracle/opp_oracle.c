/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See opp_oracle.h for the pinning statement.
 *
 * Plain-C restatement of the reference's post-processing path, one function per reference stage,
 * each citing the reference file:line it follows (paths relative to /root/reference).  Compile
 * strictly: -O2 -fno-fast-math -ffp-contract=off (x86-64 SSE2 scalar float, no FMA contraction).
 *
 * The two OpenCV calls the reference makes are restated from OpenCV's published scalar algorithms
 * (OpenCV is an external, un-vendored dependency: libopencv-dev of Ubuntu 16.04/18.04 in
 * docker/Dockerfile.builder-cpu:4; stand-in for validation: cv2 4.13 with IPP and SIMD dispatch off).
 */
#include "opp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <stddef.h>
#include <string.h>

/* include/openpose-plus/coco.h:11-53 */
static const int COCO_PAIR[ORC_N_PAIRS][2] = {
    {1, 2}, {1, 5},  {2, 3},   {3, 4},   {5, 6},  {6, 7},   {1, 8},  {8, 9},   {9, 10}, {1, 11},
    {11, 12}, {12, 13}, {1, 0}, {0, 14}, {14, 16}, {0, 15}, {15, 17}, {2, 16}, {5, 17}};
static const int COCO_NET[ORC_N_PAIRS][2] = {
    {12, 13}, {20, 21}, {14, 15}, {16, 17}, {22, 23}, {24, 25}, {0, 1},   {2, 3},   {4, 5},  {6, 7},
    {8, 9},   {10, 11}, {28, 29}, {30, 31}, {34, 35}, {32, 33}, {36, 37}, {18, 19}, {26, 27}};

void orc_coco_pair(int pair_id, int *a, int *b, int *nx, int *ny)
{
    *a = COCO_PAIR[pair_id][0];
    *b = COCO_PAIR[pair_id][1];
    *nx = COCO_NET[pair_id][0];
    *ny = COCO_NET[pair_id][1];
}

/* src/paf.cpp:60-65 -- `const float X = 0.05` members, so the literals round to float */
static const float THRESH_HEAT = 0.05f;
static const float THRESH_VECTOR_SCORE = 0.05f;
static const int THRESH_VECTOR_CNT1 = 8;
static const int THRESH_PART_CNT = 4;
static const float THRESH_HUMAN_SCORE = 0.4f;
#define STEP_PAF 10

/* ------------------------------------------------------------------------------------------- */
/* cv::getGaussianKernel(k, sigma, CV_32F): double-precision taps normalised then rounded.       */
/* Equals cv2.getGaussianKernel for every odd k in 1..63 at sigma=3 (tests/test_oracle_cv.py).   */
int orc_gauss_kernel(int ksize, double sigma, float *taps)
{
    if (ksize < 1 || (ksize & 1) == 0 || sigma <= 0) return -1;
    double t[256];
    if (ksize > 255) return -1;
    const double scale2x = -0.5 / (sigma * sigma);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) * 0.5;
        t[i] = exp(scale2x * x * x);
        sum += t[i];
    }
    const double inv = 1.0 / sum;
    for (int i = 0; i < ksize; ++i) taps[i] = (float)(t[i] * inv);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* cv::resize(..., INTER_AREA) when dst >= src on both axes (src/post-process.h:46-47).           */
/* OpenCV emulates area up-sampling by its 2-tap linear resizer with "area mode" coefficients:    */
/*   s = floor(d*scale), f = (float)((d+1) - (s+1)*inv_scale), f = f<=0 ? 0 : f - floor(f).       */
/* Returns dmax = first d whose right tap would fall outside the source (single-tap from there).  */
int orc_resize_coeffs(int ssize, int dsize, int *ofs, float *alpha)
{
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1. / inv_scale;
    int dmax = dsize;
    for (int d = 0; d < dsize; ++d) {
        int s = (int)floor(d * scale);
        float f = (float)((d + 1) - (s + 1) * inv_scale);
        f = f <= 0 ? 0.f : f - (float)(int)floor(f);
        if (s + 1 >= ssize) {
            if (d < dmax) dmax = d;
            if (s >= ssize - 1) {
                f = 0;
                s = ssize - 1;
            }
        }
        ofs[d] = s;
        alpha[2 * d] = 1.f - f;
        alpha[2 * d + 1] = f;
    }
    return dmax;
}

/* Vertical coefficients: OpenCV does not clamp fy; it clips the two row indices instead. */
static void resize_coeffs_y(int ssize, int dsize, int *ofs, float *beta)
{
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        int s = (int)floor(d * scale);
        float f = (float)((d + 1) - (s + 1) * inv_scale);
        f = f <= 0 ? 0.f : f - (float)(int)floor(f);
        ofs[d] = s;
        beta[2 * d] = 1.f - f;
        beta[2 * d + 1] = f;
    }
}

static int clip_row(int y, int n) { return y < 0 ? 0 : (y >= n ? n - 1 : y); }

typedef struct {
    int h, w, H, W;
    int *xofs, *yofs;
    float *alpha, *beta;
    int xmax;
} resize_plan;

static int plan_init(resize_plan *p, int h, int w, int H, int W)
{
    if (H < h || W < w || h < 1 || w < 1) return -1;
    p->h = h, p->w = w, p->H = H, p->W = W;
    p->xofs = (int *)malloc(sizeof(int) * W);
    p->yofs = (int *)malloc(sizeof(int) * H);
    p->alpha = (float *)malloc(sizeof(float) * 2 * W);
    p->beta = (float *)malloc(sizeof(float) * 2 * H);
    p->xmax = orc_resize_coeffs(w, W, p->xofs, p->alpha);
    resize_coeffs_y(h, H, p->yofs, p->beta);
    return 0;
}
static void plan_free(resize_plan *p)
{
    free(p->xofs), free(p->yofs), free(p->alpha), free(p->beta);
}

/* horizontal pass of one source row: D = S[sx]*a0 + S[sx+1]*a1 ; single tap from xmax on */
static void hresize_row(const resize_plan *p, const float *S, float *D)
{
    int dx = 0;
    for (; dx < p->xmax; ++dx) {
        const int sx = p->xofs[dx];
        D[dx] = S[sx] * p->alpha[2 * dx] + S[sx + 1] * p->alpha[2 * dx + 1];
    }
    for (; dx < p->W; ++dx) D[dx] = S[p->xofs[dx]] * 1.f;
}

/* one output sample, same arithmetic as the row-buffer formulation above */
static float resize_sample(const resize_plan *p, const float *src, int dy, int dx)
{
    const int sy = p->yofs[dy];
    const float *S0 = src + (size_t)clip_row(sy, p->h) * p->w;
    const float *S1 = src + (size_t)clip_row(sy + 1, p->h) * p->w;
    float r0, r1;
    const int sx = p->xofs[dx];
    if (dx < p->xmax) {
        r0 = S0[sx] * p->alpha[2 * dx] + S0[sx + 1] * p->alpha[2 * dx + 1];
        r1 = S1[sx] * p->alpha[2 * dx] + S1[sx + 1] * p->alpha[2 * dx + 1];
    } else {
        r0 = S0[sx] * 1.f;
        r1 = S1[sx] * 1.f;
    }
    return r0 * p->beta[2 * dy] + r1 * p->beta[2 * dy + 1];
}

static void resize_plane(const resize_plan *p, const float *src, float *dst, float *rows /* [h][W] */)
{
    /* OpenCV resizes the two source rows an output row needs and keeps them in a ring; resizing
       every source row once up front is the same arithmetic. */
    for (int sy = 0; sy < p->h; ++sy) hresize_row(p, src + (size_t)sy * p->w, rows + (size_t)sy * p->W);
    for (int dy = 0; dy < p->H; ++dy) {
        const int sy = p->yofs[dy];
        const float *R0 = rows + (size_t)clip_row(sy, p->h) * p->W;
        const float *R1 = rows + (size_t)clip_row(sy + 1, p->h) * p->W;
        const float b0 = p->beta[2 * dy], b1 = p->beta[2 * dy + 1];
        float *D = dst + (size_t)dy * p->W;
        for (int x = 0; x < p->W; ++x) D[x] = R0[x] * b0 + R1[x] * b1;
    }
}

int orc_resize_area_up(const float *src, int h, int w, float *dst, int H, int W)
{
    resize_plan p;
    if (plan_init(&p, h, w, H, W)) return -1;
    float *rows = (float *)malloc(sizeof(float) * (size_t)h * W);
    resize_plane(&p, src, dst, rows);
    free(rows);
    plan_free(&p);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* cv::GaussianBlur(src, dst, Size(k,k), sigma) on CV_32F, default border (REFLECT_101)          */
/* (src/post-process.h:69-70).  Separable filter, scalar op order of OpenCV's RowFilter /         */
/* SymmRowSmallFilter (k<=5) and SymmColumnFilter.                                                */
static int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0)
            p = -p;
        else
            p = 2 * (n - 1) - p;
    }
    return p;
}

int orc_gauss_blur(const float *src, int H, int W, int ksize, double sigma, float *dst)
{
    if (ksize == 1) { /* cv::GaussianBlur short-circuits a 1x1 kernel to a copy */
        memcpy(dst, src, sizeof(float) * (size_t)H * W);
        return 0;
    }
    float kx[256];
    if (orc_gauss_kernel(ksize, sigma, kx)) return -1;
    const int r = ksize / 2;
    float *row = (float *)malloc(sizeof(float) * (size_t)(W + 2 * r));
    float *tmp = (float *)malloc(sizeof(float) * (size_t)H * W);
    for (int y = 0; y < H; ++y) {
        const float *S = src + (size_t)y * W;
        for (int x = -r; x < W + r; ++x) row[x + r] = S[reflect101(x, W)];
        float *D = tmp + (size_t)y * W;
        if (ksize == 3) {
            for (int x = 0; x < W; ++x) {
                const float *p = row + x + r;
                D[x] = p[0] * kx[r] + (p[-1] + p[1]) * kx[r + 1];
            }
        } else if (ksize == 5) {
            for (int x = 0; x < W; ++x) {
                const float *p = row + x + r;
                D[x] = p[0] * kx[r] + (p[-1] + p[1]) * kx[r + 1] + (p[-2] + p[2]) * kx[r + 2];
            }
        } else {
            for (int x = 0; x < W; ++x) {
                const float *p = row + x;
                float s = kx[0] * p[0];
                for (int j = 1; j < ksize; ++j) s += kx[j] * p[j];
                D[x] = s;
            }
        }
    }
    for (int y = 0; y < H; ++y) {
        float *D = dst + (size_t)y * W;
        const float *c = tmp + (size_t)y * W;
        for (int x = 0; x < W; ++x) D[x] = kx[r] * c[x] + 0.f;
        for (int j = 1; j <= r; ++j) {
            const float *a = tmp + (size_t)reflect101(y + j, H) * W;
            const float *b = tmp + (size_t)reflect101(y - j, H) * W;
            const float f = kx[r + j];
            for (int x = 0; x < W; ++x) D[x] += f * (a[x] + b[x]);
        }
    }
    free(row);
    free(tmp);
    return 0;
}

/* Python-path variant.  openpose_plus/inference/post_process.py:13-17: _gauss_kernel(ksize, nsig) =
 * sqrt(outer(y, y)) / sum, y = diff(norm.cdf(linspace(-nsig - i/2, nsig + i/2, ksize + 1))), i = (2 nsig + 1) / ksize.
 * sqrt(outer(y, y)) = outer(sqrt y, sqrt y), so the 2-D filter is the outer product of g = sqrt(y) / sum(sqrt(y)). */
int orc_cdf_kernel(int ksize, double nsig, float *taps)
{
    if (ksize < 1 || ksize > 255) return -1;
    double e[257], t[256], sum = 0;
    const double interval = (2 * nsig + 1.) / ksize, lo = -nsig - interval / 2., hi = nsig + interval / 2.;
    for (int i = 0; i <= ksize; ++i) {
        const double x = i == ksize ? hi : lo + (hi - lo) / ksize * i;
        e[i] = 0.5 * erfc(-x / sqrt(2.0));
    }
    for (int i = 0; i < ksize; ++i) t[i] = sqrt(e[i + 1] - e[i]), sum += t[i];
    for (int i = 0; i < ksize; ++i) taps[i] = (float)(t[i] / sum);
    return 0;
}

/* tf.nn.depthwise_conv2d(..., padding='SAME') with that filter (post_process.py:20-26): pixels outside the image
 * are zero.  TensorFlow fixes no summation order; this restatement (and the CUDA kernel) evaluates the separable
 * form in float32: row pass left to right, column pass centre first then symmetric pairs outwards. */
int orc_smooth_zero_pad(const float *src, int H, int W, int ksize, const float *k, float *dst)
{
    if (ksize < 1 || !(ksize & 1)) return -1;
    const int r = ksize / 2;
    float *row = (float *)calloc((size_t)(W + 2 * r), sizeof(float));
    float *tmp = (float *)calloc((size_t)(H + 2 * r) * W, sizeof(float)); /* rows -r .. H+r-1, zero outside */
    for (int y = 0; y < H; ++y) {
        memcpy(row + r, src + (size_t)y * W, sizeof(float) * (size_t)W);
        float *D = tmp + (size_t)(y + r) * W;
        for (int x = 0; x < W; ++x) {
            const float *p = row + x;
            float s = k[0] * p[0];
            for (int j = 1; j < ksize; ++j) s += k[j] * p[j];
            D[x] = s;
        }
    }
    for (int y = 0; y < H; ++y) {
        float *D = dst + (size_t)y * W;
        const float *c = tmp + (size_t)(y + r) * W;
        for (int x = 0; x < W; ++x) {
            float s = k[r] * c[x];
            for (int j = 1; j <= r; ++j) s += k[r + j] * (c[x + (ptrdiff_t)j * W] + c[x - (ptrdiff_t)j * W]);
            D[x] = s;
        }
    }
    free(row), free(tmp);
    return 0;
}

/* src/post-process.h:74-95 (CPU) == cuDNN 3x3/stride 1/pad 1 max pool for non-NaN input */
void orc_max_pool_3x3(const float *src, int H, int W, float *dst)
{
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float m = src[(size_t)i * W + j];
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj) {
                    const int y = i + di, x = j + dj;
                    if (0 <= y && y < H && 0 <= x && x < W) {
                        const float v = src[(size_t)y * W + x];
                        if (m < v) m = v; /* std::max(a,b) = (a<b)?b:a */
                    }
                }
            dst[(size_t)i * W + j] = m;
        }
}

/* ------------------------------------------------------------------------------------------- */
/* std::sort(first,last,std::greater<>) exactly as libstdc++ (GCC 13, bits/stl_algo.h:1942 ff.)  */
/* runs it for src/paf.cpp:151-152.  Element movement decides the order of equal scores.         */
#define GT(a, b) ((a).score > (b).score)

static void swap_c(orc_cand_t *a, orc_cand_t *b)
{
    orc_cand_t t = *a;
    *a = *b;
    *b = t;
}

static void move_median_to_first(orc_cand_t *result, orc_cand_t *a, orc_cand_t *b, orc_cand_t *c)
{
    if (GT(*a, *b)) {
        if (GT(*b, *c))
            swap_c(result, b);
        else if (GT(*a, *c))
            swap_c(result, c);
        else
            swap_c(result, a);
    } else if (GT(*a, *c))
        swap_c(result, a);
    else if (GT(*b, *c))
        swap_c(result, c);
    else
        swap_c(result, b);
}

static orc_cand_t *unguarded_partition(orc_cand_t *first, orc_cand_t *last, orc_cand_t *pivot)
{
    for (;;) {
        while (GT(*first, *pivot)) ++first;
        --last;
        while (GT(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swap_c(first, last);
        ++first;
    }
}

static void push_heap_(orc_cand_t *first, long hole, long top, orc_cand_t value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && GT(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void adjust_heap(orc_cand_t *first, long hole, long len, orc_cand_t value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (GT(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value);
}

static void heap_sort_all(orc_cand_t *first, orc_cand_t *last)
{
    /* __partial_sort(first,last,last): __heap_select with middle==last is just make_heap */
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            orc_cand_t v = first[parent];
            adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        orc_cand_t v = *last;
        *last = *first;
        adjust_heap(first, 0, last - first, v);
    }
}

static void introsort_loop(orc_cand_t *first, orc_cand_t *last, long depth_limit)
{
    while (last - first > 16) {
        if (depth_limit == 0) {
            heap_sort_all(first, last);
            return;
        }
        --depth_limit;
        orc_cand_t *mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        orc_cand_t *cut = unguarded_partition(first + 1, last, first);
        introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}

static void unguarded_linear_insert(orc_cand_t *last)
{
    orc_cand_t val = *last;
    orc_cand_t *next = last - 1;
    while (GT(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

static void insertion_sort(orc_cand_t *first, orc_cand_t *last)
{
    if (first == last) return;
    for (orc_cand_t *i = first + 1; i != last; ++i) {
        if (GT(*i, *first)) {
            orc_cand_t val = *i;
            memmove(first + 1, first, sizeof(orc_cand_t) * (size_t)(i - first));
            *first = val;
        } else
            unguarded_linear_insert(i);
    }
}

void orc_std_sort_desc(orc_cand_t *v, int n)
{
    if (n <= 0) return;
    long lg = 0;
    for (long t = n; t > 1; t >>= 1) ++lg;
    introsort_loop(v, v + n, 2 * lg);
    if (n > 16) {
        insertion_sort(v, v + 16);
        for (orc_cand_t *i = v + 16; i != v + n; ++i) unguarded_linear_insert(i);
    } else
        insertion_sort(v, v + n);
}

/* ------------------------------------------------------------------------------------------- */
struct orc_ctx {
    int h, w, H, W, ksize;
    int variant; /* 0 = C++ path (src/paf.cpp, parity target), 1 = Python path (post_process.py + pafprocess) */
    resize_plan plan;
    float *rows;
    float *conf_up, *paf_up, *smoothed, *pooled;
    const float *paf_lowres; /* lazy mode */
    int lazy;
    orc_peak_t *peaks;
    int n_peaks, cap_peaks;
    int part_ofs[ORC_N_PARTS + 1];
    orc_cand_t *cands_raw[ORC_N_PAIRS], *cands_sorted[ORC_N_PAIRS];
    int n_cands[ORC_N_PAIRS], cap_cands[ORC_N_PAIRS], n_pairs[ORC_N_PAIRS], ties[ORC_N_PAIRS];
    orc_conn_t *conns[ORC_N_PAIRS];
    int n_conns[ORC_N_PAIRS], cap_conns[ORC_N_PAIRS];
    orc_href_t *hrefs;
    int n_hrefs, cap_hrefs, hist_max, n_incomplete, n_merges;
    orc_human_t *humans;
    int n_humans, cap_humans;
    int flags;
};

orc_ctx *orc_create(int h, int w, int H, int W, int ksize)
{
    if (ksize < 1 || !(ksize & 1) || ksize > 255) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    if (plan_init(&c->plan, h, w, H, W)) {
        free(c);
        return NULL;
    }
    c->h = h, c->w = w, c->H = H, c->W = W, c->ksize = ksize;
    const size_t px = (size_t)H * W;
    c->rows = (float *)malloc(sizeof(float) * (size_t)h * W);
    c->conf_up = (float *)malloc(sizeof(float) * ORC_N_HEAT * px);
    c->paf_up = (float *)malloc(sizeof(float) * ORC_N_PAF * px);
    c->smoothed = (float *)malloc(sizeof(float) * ORC_N_HEAT * px);
    c->pooled = (float *)malloc(sizeof(float) * ORC_N_HEAT * px);
    return c;
}

void orc_set_variant(orc_ctx *c, int variant) { c->variant = variant; }

void orc_destroy(orc_ctx *c)
{
    if (!c) return;
    plan_free(&c->plan);
    free(c->rows), free(c->conf_up), free(c->paf_up), free(c->smoothed), free(c->pooled);
    free(c->peaks), free(c->hrefs), free(c->humans);
    for (int i = 0; i < ORC_N_PAIRS; ++i) free(c->cands_raw[i]), free(c->cands_sorted[i]), free(c->conns[i]);
    free(c);
}

/* src/post-process.h:176-199: raster scan k -> i -> j; the running index is the peak id */
static void find_peaks(orc_ctx *c)
{
    const size_t px = (size_t)c->H * c->W;
    c->n_peaks = 0;
    size_t off = 0;
    for (int k = 0; k < ORC_N_HEAT; ++k) {
        if (k < ORC_N_PARTS) c->part_ofs[k] = c->n_peaks;
        for (int i = 0; i < c->H; ++i)
            for (int j = 0; j < c->W; ++j, ++off) {
                if (k < ORC_N_PARTS && c->smoothed[off] > THRESH_HEAT && c->smoothed[off] == c->pooled[off]) {
                    if (c->n_peaks == c->cap_peaks) {
                        c->cap_peaks = c->cap_peaks ? 2 * c->cap_peaks : 1024;
                        c->peaks = (orc_peak_t *)realloc(c->peaks, sizeof(orc_peak_t) * c->cap_peaks);
                    }
                    orc_peak_t *p = &c->peaks[c->n_peaks];
                    p->part_id = k, p->x = j, p->y = i, p->score = c->conf_up[off], p->id = c->n_peaks;
                    c->n_peaks++;
                }
            }
    }
    c->part_ofs[ORC_N_PARTS] = c->n_peaks;
    (void)px;
}

static float paf_at(const orc_ctx *c, int ch, int y, int x)
{
    if (c->lazy) return resize_sample(&c->plan, c->paf_lowres + (size_t)ch * c->h * c->w, y, x);
    return c->paf_up[((size_t)ch * c->H + y) * c->W + x];
}

/* src/paf.cpp:337 -- float argument, double add, truncation */
static int roundpaf(float v) { return (int)(v + 0.5); }

/* src/paf.cpp:79-134 with get_paf_vectors :313-335 folded in */
static void score_pair_list(orc_ctx *c, int pair_id)
{
    const int pa = COCO_PAIR[pair_id][0], pb = COCO_PAIR[pair_id][1];
    const int cx = COCO_NET[pair_id][0], cy = COCO_NET[pair_id][1];
    const int height = c->H; /* src/paf.cpp:274 passes the up-sampled height */
    c->n_cands[pair_id] = 0;
    c->n_pairs[pair_id] = 0;
    for (int ia = c->part_ofs[pa]; ia < c->part_ofs[pa + 1]; ++ia)
        for (int ib = c->part_ofs[pb]; ib < c->part_ofs[pb + 1]; ++ib) {
            const orc_peak_t *A = &c->peaks[ia], *B = &c->peaks[ib];
            c->n_pairs[pair_id]++;
            const int dx = B->x - A->x, dy = B->y - A->y;
            const float norm = (float)sqrt((double)(dx * dx + dy * dy));
            if (norm < 1e-12) continue;
            float vx = (float)dx, vy = (float)dy;
            vx /= norm;
            vy /= norm;
            const float step_x = (float)dx / (float)STEP_PAF;
            const float step_y = (float)dy / (float)STEP_PAF;
            float scores = 0.0f;
            int criterion1 = 0;
            for (int i = 0; i < STEP_PAF; ++i) {
                const int lx = roundpaf((float)A->x + (float)i * step_x);
                const int ly = roundpaf((float)A->y + (float)i * step_y);
                const float px = paf_at(c, cx, ly, lx);
                const float py = paf_at(c, cy, ly, lx);
                const float score = vx * px + vy * py;
                scores += score;
                if (score > THRESH_VECTOR_SCORE) criterion1 += 1;
            }
            const double pen = 0.5 * height / norm - 1.0;
            const float criterion2 = (float)((double)(scores / (float)STEP_PAF) + (pen < 0.0 ? pen : 0.0)); /* std::min(0.0, pen) */
            if (criterion1 > THRESH_VECTOR_CNT1 && criterion2 > 0) {
                if (c->n_cands[pair_id] == c->cap_cands[pair_id]) {
                    const int cap = c->cap_cands[pair_id] ? 2 * c->cap_cands[pair_id] : 256;
                    c->cap_cands[pair_id] = cap;
                    c->cands_raw[pair_id] = (orc_cand_t *)realloc(c->cands_raw[pair_id], sizeof(orc_cand_t) * cap);
                    c->cands_sorted[pair_id] = (orc_cand_t *)realloc(c->cands_sorted[pair_id], sizeof(orc_cand_t) * cap);
                }
                orc_cand_t *cd = &c->cands_raw[pair_id][c->n_cands[pair_id]++];
                cd->idx1 = A->id, cd->idx2 = B->id, cd->score = criterion2;
                cd->etc = criterion2 + A->score + B->score;
            }
        }
}

/* src/paf.cpp:136-175 */
static void match_pair_list(orc_ctx *c, int pair_id)
{
    const int n = c->n_cands[pair_id];
    if (n) memcpy(c->cands_sorted[pair_id], c->cands_raw[pair_id], sizeof(orc_cand_t) * n);
    orc_cand_t *v = c->cands_sorted[pair_id];
    orc_std_sort_desc(v, n);
    c->ties[pair_id] = 0;
    for (int i = 1; i < n; ++i)
        if (v[i].score == v[i - 1].score) c->ties[pair_id] = 1;
    c->n_conns[pair_id] = 0;
    for (int i = 0; i < n; ++i) {
        int assigned = 0;
        for (int k = 0; k < c->n_conns[pair_id]; ++k)
            if (c->conns[pair_id][k].cid1 == v[i].idx1 || c->conns[pair_id][k].cid2 == v[i].idx2) {
                assigned = 1;
                break;
            }
        if (!assigned) {
            if (c->n_conns[pair_id] == c->cap_conns[pair_id]) {
                const int cap = c->cap_conns[pair_id] ? 2 * c->cap_conns[pair_id] : 64;
                c->cap_conns[pair_id] = cap;
                c->conns[pair_id] = (orc_conn_t *)realloc(c->conns[pair_id], sizeof(orc_conn_t) * cap);
            }
            orc_conn_t *cn = &c->conns[pair_id][c->n_conns[pair_id]++];
            cn->cid1 = v[i].idx1, cn->cid2 = v[i].idx2, cn->score = v[i].score;
        }
    }
}

static float peak_score(orc_ctx *c, int id)
{
    if (id < 0 || id >= c->n_peaks) {
        c->flags |= ORC_FLAG_UB_PEAK_INDEX;
        return 0.f;
    }
    return c->peaks[id].score;
}

/* human_refs[id] with std::vector storage semantics: slots in [size, hist_max) still hold the
 * bytes left behind by erase()'s shift-down; anything at or beyond hist_max was never written. */
static orc_href_t *href_at(orc_ctx *c, int idx)
{
    static orc_href_t dummy;
    if (idx < 0 || idx >= c->hist_max) {
        c->flags |= ORC_FLAG_UB_STALE_INDEX;
        memset(&dummy, 0xff, sizeof dummy);
        dummy.score = 0, dummy.n_parts = 0;
        return &dummy;
    }
    return &c->hrefs[idx];
}

/* src/paf.cpp:177-262, bugs included (stale hr.id used as an index; `> 0` membership test) */
static void assemble(orc_ctx *c)
{
    c->n_hrefs = 0, c->hist_max = 0, c->n_merges = 0;
    for (int pair_id = 0; pair_id < ORC_N_PAIRS; ++pair_id) {
        const int part1 = COCO_PAIR[pair_id][0], part2 = COCO_PAIR[pair_id][1];
        for (int k = 0; k < c->n_conns[pair_id]; ++k) {
            const orc_conn_t conn = c->conns[pair_id][k];
            int hits[2], n_hits = 0;
            for (int q = 0; q < c->n_hrefs; ++q) {
                const orc_href_t *hr = &c->hrefs[q];
                if (hr->parts[part1] == conn.cid1 || hr->parts[part2] == conn.cid2) {
                    /* src/paf.cpp:198 records the STORED id; pafprocess (Python path) records the position */
                    if (n_hits < 2) hits[n_hits] = c->variant == 1 ? q : hr->id;
                    n_hits++;
                }
            }
            if (n_hits == 1) {
                orc_href_t *hr1 = href_at(c, hits[0]);
                if (hr1->parts[part2] != conn.cid2) {
                    hr1->parts[part2] = conn.cid2;
                    ++hr1->n_parts;
                    hr1->score += peak_score(c, conn.cid2) + conn.score;
                }
            } else if (n_hits >= 2) {
                orc_href_t *hr1 = href_at(c, hits[0]);
                orc_href_t *hr2 = href_at(c, hits[1]);
                int membership = 0;
                for (int i = 0; i < ORC_N_PARTS; ++i)
                    if (hr1->parts[i] > 0 && hr2->parts[i] > 0) membership = 2;
                if (membership == 0) {
                    for (int i = 0; i < ORC_N_PARTS; ++i) hr1->parts[i] += hr2->parts[i] + 1;
                    hr1->n_parts += hr2->n_parts;
                    hr1->score += hr2->score;
                    hr1->score += conn.score;
                    /* human_refs.erase(begin() + hr_ids[1]) */
                    const int e = hits[1];
                    if (e < 0 || e >= c->n_hrefs) {
                        /* erase() at or past end(): undefined by the standard.  libstdc++ 13's
                           vector::_M_erase moves nothing (negative distance) and still pops the last
                           element; the strict build of the reference made here behaves that way, so
                           it is restated and the frame is flagged. */
                        c->flags |= ORC_FLAG_UB_ERASE_PAST_END;
                        if (c->n_hrefs > 0) c->n_hrefs--;
                    } else {
                        memmove(&c->hrefs[e], &c->hrefs[e + 1], sizeof(orc_href_t) * (size_t)(c->n_hrefs - e - 1));
                        c->n_hrefs--;
                    }
                    c->n_merges++;
                } else {
                    hr1->parts[part2] = conn.cid2;
                    hr1->n_parts += 1;
                    hr1->score += peak_score(c, conn.cid2) + conn.score;
                }
            } else if (n_hits == 0 && !(pair_id > 16)) {
                if (c->n_hrefs == c->cap_hrefs) {
                    c->cap_hrefs = c->cap_hrefs ? 2 * c->cap_hrefs : 64;
                    c->hrefs = (orc_href_t *)realloc(c->hrefs, sizeof(orc_href_t) * c->cap_hrefs);
                }
                orc_href_t *hnew = &c->hrefs[c->n_hrefs];
                hnew->id = c->n_hrefs;
                for (int i = 0; i < ORC_N_PARTS; ++i) hnew->parts[i] = -1;
                hnew->parts[part1] = conn.cid1;
                hnew->parts[part2] = conn.cid2;
                hnew->n_parts = 2;
                hnew->score = peak_score(c, conn.cid1) + peak_score(c, conn.cid2) + conn.score;
                c->n_hrefs++;
                if (c->n_hrefs > c->hist_max) c->hist_max = c->n_hrefs;
            }
        }
    }
    c->n_incomplete = c->n_hrefs;
    /* src/paf.cpp:253-260 std::remove_if keeps relative order */
    int out = 0;
    for (int q = 0; q < c->n_hrefs; ++q) {
        const orc_href_t hr = c->hrefs[q];
        if (hr.n_parts < THRESH_PART_CNT || hr.score / hr.n_parts < THRESH_HUMAN_SCORE) continue;
        c->hrefs[out++] = hr;
    }
    c->n_hrefs = out;
}

/* src/paf.cpp:292-310 */
static void emit_humans(orc_ctx *c)
{
    if (c->cap_humans < c->n_hrefs) {
        c->cap_humans = c->n_hrefs + 16;
        c->humans = (orc_human_t *)realloc(c->humans, sizeof(orc_human_t) * c->cap_humans);
    }
    c->n_humans = c->n_hrefs;
    for (int q = 0; q < c->n_hrefs; ++q) {
        orc_human_t *h = &c->humans[q];
        memset(h, 0, sizeof *h);
        h->score = c->hrefs[q].score;
        for (int i = 0; i < ORC_N_PARTS; ++i) {
            const int id = c->hrefs[q].parts[i];
            if (id != -1) {
                h->parts[i].has_value = 1;
                if (id < 0 || id >= c->n_peaks) {
                    c->flags |= ORC_FLAG_UB_PEAK_INDEX;
                    continue;
                }
                h->parts[i].score = c->peaks[id].score;
                h->parts[i].x = (float)c->peaks[id].x;
                h->parts[i].y = (float)c->peaks[id].y;
            }
        }
    }
}

static int run(orc_ctx *c, const float *conf, const float *paf, int lazy)
{
    const size_t lp = (size_t)c->h * c->w, px = (size_t)c->H * c->W;
    c->flags = 0;
    c->lazy = lazy;
    c->paf_lowres = paf;
    /* src/paf.cpp:46-51 */
    for (int k = 0; k < ORC_N_HEAT; ++k) resize_plane(&c->plan, conf + k * lp, c->conf_up + k * px, c->rows);
    if (!lazy)
        for (int k = 0; k < ORC_N_PAF; ++k) resize_plane(&c->plan, paf + k * lp, c->paf_up + k * px, c->rows);
    /* src/post-process.h:160-174 (use_gpu=false branch; the cuDNN branch computes the same max) */
    for (int k = 0; k < ORC_N_HEAT; ++k) {
        if (c->variant == 1) {
            float taps[256];
            orc_cdf_kernel(c->ksize, 3.0, taps);
            orc_smooth_zero_pad(c->conf_up + k * px, c->H, c->W, c->ksize, taps, c->smoothed + k * px);
        } else
            orc_gauss_blur(c->conf_up + k * px, c->H, c->W, c->ksize, 3.0, c->smoothed + k * px);
        orc_max_pool_3x3(c->smoothed + k * px, c->H, c->W, c->pooled + k * px);
    }
    find_peaks(c);
    for (int p = 0; p < ORC_N_PAIRS; ++p) {
        score_pair_list(c, p);
        match_pair_list(c, p);
    }
    assemble(c);
    emit_humans(c);
    return c->n_humans;
}

int orc_run(orc_ctx *c, const float *conf, const float *paf) { return run(c, conf, paf, 0); }
int orc_run_lazy(orc_ctx *c, const float *conf, const float *paf) { return run(c, conf, paf, 1); }

const float *orc_conf_up(const orc_ctx *c) { return c->conf_up; }
const float *orc_paf_up(const orc_ctx *c) { return c->paf_up; }
const float *orc_smoothed(const orc_ctx *c) { return c->smoothed; }
const float *orc_pooled(const orc_ctx *c) { return c->pooled; }
int orc_n_peaks(const orc_ctx *c) { return c->n_peaks; }
const orc_peak_t *orc_peaks(const orc_ctx *c) { return c->peaks; }
int orc_n_pairs_scored(const orc_ctx *c, int p) { return c->n_pairs[p]; }
int orc_n_cands(const orc_ctx *c, int p) { return c->n_cands[p]; }
const orc_cand_t *orc_cands_unsorted(const orc_ctx *c, int p) { return c->cands_raw[p]; }
const orc_cand_t *orc_cands_sorted(const orc_ctx *c, int p) { return c->cands_sorted[p]; }
int orc_has_score_ties(const orc_ctx *c, int p) { return c->ties[p]; }
int orc_n_conns(const orc_ctx *c, int p) { return c->n_conns[p]; }
const orc_conn_t *orc_conns(const orc_ctx *c, int p) { return c->conns[p]; }
int orc_n_incomplete(const orc_ctx *c) { return c->n_incomplete; }
int orc_n_merges(const orc_ctx *c) { return c->n_merges; }
int orc_n_humans(const orc_ctx *c) { return c->n_humans; }
const orc_href_t *orc_hrefs(const orc_ctx *c) { return c->hrefs; }
const orc_human_t *orc_humans(const orc_ctx *c) { return c->humans; }
int orc_flags(const orc_ctx *c) { return c->flags; }
