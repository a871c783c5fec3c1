/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the openpose-plus post-processing hot path, with every
 * intermediate exposed so the CUDA path can be checked stage by stage.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity pinning: the reference holds no golden vectors for this path (CMakeLists.txt:24-25,
 * "# TODO: add tests").  This restatement is instead pinned against
 *   (1) oracle/_ref/libopp_ref.so = the reference's own, unmodified src/paf.cpp compiled here
 *       (oracle/Makefile) -- final human_t lists must be identical, and
 *   (2) cv2 4.13 (IPP off, setUseOptimized(False)) for the two OpenCV calls the reference makes
 *       (cv::resize INTER_AREA, cv::GaussianBlur) -- bit-identical, see tests/test_oracle_cv.py
 * and the resulting vectors are committed under tests/golden/.
 */
#ifndef OPP_ORACLE_H
#define OPP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_N_PARTS 18
#define ORC_N_PAIRS 19
#define ORC_N_HEAT 19
#define ORC_N_PAF 38

/* src/post-process.h:132-137 */
typedef struct {
    int part_id;
    int x, y;
    float score;
    int id;
} orc_peak_t;

/* include/openpose-plus/human.h:36-41 */
typedef struct {
    int idx1, idx2;
    float score, etc;
} orc_cand_t;

/* include/openpose-plus/human.h:49-55 (cidN == peak_idN always, src/paf.cpp:166-170) */
typedef struct {
    int cid1, cid2;
    float score;
} orc_conn_t;

/* include/openpose-plus/human.h:57-77 */
typedef struct {
    int id;
    int parts[ORC_N_PARTS];
    float score;
    int n_parts;
} orc_href_t;

/* include/openpose-plus/human.h:8-34; has_value is a bool followed by 3 pad bytes */
typedef struct {
    unsigned char has_value;
    unsigned char pad_[3];
    float x, y, score;
} orc_part_t;
typedef struct {
    orc_part_t parts[ORC_N_PARTS];
    float score;
} orc_human_t;

/* flag bits of orc_flags(): frames whose reference behaviour is undefined */
#define ORC_FLAG_UB_STALE_INDEX 1 /* human_refs[] indexed at or beyond its historical maximum size */
#define ORC_FLAG_UB_PEAK_INDEX 2  /* all_peaks[] indexed with a corrupted (merged) id */
#define ORC_FLAG_UB_ERASE_PAST_END 4 /* human_refs.erase() with a stale id >= size(): restated as libstdc++ 13 behaves */

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(int feat_h, int feat_w, int out_h, int out_w, int ksize);
void orc_destroy(orc_ctx *);
/* 0 (default) = the C++ path, src/paf.cpp -- the parity target, pinned against the reference build.
 * 1 = the Python path's semantics (openpose_plus/inference/post_process.py:13-37 smoothing with the CDF-derived
 *     kernel and zero padding; grouping as the external tf_pose pafprocess module: humans indexed by position).
 *     PARITY UNPINNED for this variant: TensorFlow and pafprocess are absent from the reference tree and from this
 *     image; the smoothing is checked against a float64 2-D convolution within tolerance only. */
void orc_set_variant(orc_ctx *, int variant);
int orc_cdf_kernel(int ksize, double nsig, float *taps);
int orc_smooth_zero_pad(const float *src, int H, int W, int ksize, const float *taps, float *dst);
/* Runs the whole path on one frame (conf [19,h,w], paf [38,h,w]); returns the number of humans. */
int orc_run(orc_ctx *, const float *conf, const float *paf);
/* Same, but skips materialising paf_up (samples are computed on demand, bit-identical). */
int orc_run_lazy(orc_ctx *, const float *conf, const float *paf);

const float *orc_conf_up(const orc_ctx *);  /* [19,H,W] */
const float *orc_paf_up(const orc_ctx *);   /* [38,H,W] */
const float *orc_smoothed(const orc_ctx *); /* [19,H,W] */
const float *orc_pooled(const orc_ctx *);   /* [19,H,W] */
int orc_n_peaks(const orc_ctx *);
const orc_peak_t *orc_peaks(const orc_ctx *);
int orc_n_pairs_scored(const orc_ctx *, int pair_id);
int orc_n_cands(const orc_ctx *, int pair_id);
const orc_cand_t *orc_cands_unsorted(const orc_ctx *, int pair_id);
const orc_cand_t *orc_cands_sorted(const orc_ctx *, int pair_id);
int orc_has_score_ties(const orc_ctx *, int pair_id);
int orc_n_conns(const orc_ctx *, int pair_id);
const orc_conn_t *orc_conns(const orc_ctx *, int pair_id);
int orc_n_incomplete(const orc_ctx *);
int orc_n_merges(const orc_ctx *);
int orc_n_humans(const orc_ctx *);
const orc_href_t *orc_hrefs(const orc_ctx *);   /* surviving human refs, output order */
const orc_human_t *orc_humans(const orc_ctx *); /* human_t records, output order */
int orc_flags(const orc_ctx *);

/* stand-alone pieces (also used by oracle/cv_standin.cpp) */
int orc_gauss_kernel(int ksize, double sigma, float *taps);
int orc_resize_area_up(const float *src, int h, int w, float *dst, int H, int W);
int orc_resize_coeffs(int ssize, int dsize, int *ofs, float *alpha /* [dsize][2] */);
int orc_gauss_blur(const float *src, int H, int W, int ksize, double sigma, float *dst);
void orc_max_pool_3x3(const float *src, int H, int W, float *dst);
void orc_std_sort_desc(orc_cand_t *v, int n);

/* table access for tests */
void orc_coco_pair(int pair_id, int *part_a, int *part_b, int *net_x, int *net_y);

#ifdef __cplusplus
}
#endif
#endif
