/*
 * opp_b200.h -- C-ABI of the B200-native openpose-plus post-processing path
 * (part-confidence maps + part-affinity fields in, grouped COCO-18 skeletons out).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Two host sides sit
 * on top of it: the C++ `paf_processor` subclass behind the unchanged `create_paf_processor`
 * (csrc/paf_processor.cpp; replaces /root/reference src/paf.cpp:19-57,340-346) and the Python
 * `PostProcessor` (openpose_plus_b200/post_process.py; replaces
 * openpose_plus/inference/post_process.py:109-150) through ctypes.
 *
 * Everything behind it runs as hand-written sm_100a CUDA kernels; there is no CPU fallback.
 * Reference file:line citations are relative to /root/reference.
 */
#ifndef OPP_B200_H
#define OPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPP_N_PARTS 18 /* COCO_N_PARTS, include/openpose-plus/coco.h:5 */
#define OPP_N_PAIRS 19 /* COCO_N_PAIRS, include/openpose-plus/coco.h:6 */
#define OPP_N_HEAT 19  /* n_joins,  include/openpose-plus.h:11 */
#define OPP_N_PAF 38   /* 2 * n_connections, include/openpose-plus.h:12 */

/* status codes (the reference has none: cuDNN errors exit(1), src/cudnn_traits.hpp:11-20) */
enum {
    OPP_OK = 0,
    OPP_ERR_INVALID = 1,  /* bad argument / unsupported geometry */
    OPP_ERR_CUDA = 2,     /* a CUDA runtime call failed; see opp_last_error */
    OPP_ERR_NO_DEVICE = 3,
    OPP_ERR_BUSY = 4      /* no free pipeline slot (opp_submit without opp_wait) */
};

/* per-frame flag bits written to frame_flags[] */
enum {
    OPP_FLAG_PEAK_OVERFLOW = 1,   /* a part had more peaks than max_peaks_per_part: result invalid */
    OPP_FLAG_CAND_OVERFLOW = 2,   /* a limb had more accepted candidates than max_cands_per_limb */
    OPP_FLAG_HUMAN_OVERFLOW = 4,  /* more partial humans than max_humans */
    OPP_FLAG_UB_STALE_INDEX = 8,  /* reference would index human_refs[] beyond its historical size (src/paf.cpp:204,211-212) */
    OPP_FLAG_UB_PEAK_INDEX = 16,  /* reference would index all_peaks[] with a corrupted merged id (src/paf.cpp:224,301) */
    OPP_FLAG_UB_ERASE_PAST_END = 32 /* reference erases at a stale id >= size() (src/paf.cpp:231); restated as libstdc++ 13 behaves */
};

/* Which of the reference's two post-processing semantics a handle reproduces.
 *  OPP_VARIANT_CPP    -- the C++ path, the parity target: cv::GaussianBlur(k, sigma 3) with REFLECT_101 borders
 *                        (src/post-process.h:51-72) and getHumans with its stored-id indexing (src/paf.cpp:177-262).
 *  OPP_VARIANT_PYTHON -- the Python path (openpose_plus/inference/post_process.py:13-37,82-106): separable form of
 *                        the CDF-derived kernel of _gauss_kernel(k, 3.0) with tf 'SAME' zero padding (the reference
 *                        fixes k = 25), peaks where smoothed == 3x3 max-pool and smoothed > THRESH_HEAT, grouping as
 *                        the external tf_pose `pafprocess` module does it (same scoring / sort / greedy matching as
 *                        src/paf.cpp, humans indexed by position instead of stored id).  TensorFlow fixes no summation
 *                        order and pafprocess is not vendored in the reference: parity of this variant is pinned only
 *                        to this repo's restatement (oracle/), within float tolerance of a float64 2-D convolution. */
enum { OPP_VARIANT_CPP = 0, OPP_VARIANT_PYTHON = 1 };

enum { OPP_MEM_HOST = 0, OPP_MEM_DEVICE = 1 };
enum { OPP_LAYOUT_CHW = 0, OPP_LAYOUT_HWC = 1 };
/* How DEVICE-resident inputs of a batch become ready (opp_batch_t.in_sync).  The feature maps normally come from a
 * CNN runner working on its own stream (src/uff-runner.cpp:199-205 executes the TensorRT context, then copies the maps
 * to the host only because paf_processor wants host pointers): with STREAM / EVENT the hand-off is ordered on the
 * device and the caller never synchronises the host. */
enum {
    OPP_SYNC_NONE = 0,   /* the inputs are complete when opp_submit is called (host memory, or the caller synchronised) */
    OPP_SYNC_STREAM = 1, /* in_sync_obj is the producer's cudaStream_t (NULL = the legacy default stream): the batch
                          * waits for everything enqueued on it so far (the library records its own event there) */
    OPP_SYNC_EVENT = 2   /* in_sync_obj is a cudaEvent_t the caller has recorded after the producer's last write */
};

/* include/openpose-plus/human.h:8-15 (bool + 3 pad bytes, then 3 floats) */
typedef struct {
    uint8_t has_value;
    uint8_t pad_[3];
    float x, y, score;
} opp_body_part_t;

/* include/openpose-plus/human.h:17-34 -- human_t, 292 bytes; x,y are pixels of the up-sampled map */
typedef struct {
    opp_body_part_t parts[OPP_N_PARTS];
    float score;
} opp_human_t;

/* src/post-process.h:132-137 -- peak_info */
typedef struct {
    int32_t part_id;
    int32_t x, y;
    float score;
    int32_t id;
} opp_peak_t;

/* include/openpose-plus/human.h:49-55 -- Connection (cidN == peak_idN, src/paf.cpp:166-170) */
typedef struct {
    int32_t cid1, cid2;
    float score;
} opp_conn_t;

/* Arguments of create_paf_processor (include/openpose-plus.hpp:54-64) plus capacities.  The
 * reference grows std::vectors without bound; fixed capacities are reported through frame_flags. */
typedef struct {
    int32_t feat_h, feat_w;      /* "input_height/input_width": size of the feature maps */
    int32_t out_h, out_w;        /* "height/width": size the maps are up-sampled to */
    int32_t n_joins;             /* must be 19 */
    int32_t n_connections;       /* must be 19 */
    int32_t gauss_kernel_size;   /* odd, 1..63 */
    int32_t max_batch;           /* frames per opp_process/opp_submit call */
    int32_t device;              /* CUDA ordinal, -1 = current device */
    int32_t max_peaks_per_part;  /* default 128 */
    int32_t max_cands_per_limb;  /* default 1024 */
    int32_t max_humans;          /* default 128 (counts partial humans during assembly) */
    int32_t n_slots;             /* batches in flight (own stream + buffers each), default 3 */
    int32_t variant;             /* OPP_VARIANT_CPP (default) | OPP_VARIANT_PYTHON */
    int32_t reserved[2];
} opp_config_t;

/* One batch of frames. */
typedef struct {
    const float *conf;     /* [n, 19, h, w] (CHW) or [n, h, w, 19] (HWC) */
    const float *paf;      /* [n, 38, h, w] (CHW) or [n, h, w, 38] (HWC) */
    int32_t n_frames;      /* 1..max_batch */
    int32_t in_mem;        /* OPP_MEM_HOST | OPP_MEM_DEVICE */
    int32_t in_layout;     /* OPP_LAYOUT_CHW | OPP_LAYOUT_HWC */
    int32_t out_mem;       /* where humans / n_humans / frame_flags live */
    opp_human_t *humans;   /* [n, max_humans]; frame f's humans are humans[f*max_humans .. +n_humans[f]).  Pinned host
                            * buffers (opp_host_alloc / cudaHostRegister) are written directly by the GPU, pageable ones
                            * through the slot's pinned staging; entries beyond n_humans[f] are left untouched */
    int32_t *n_humans;     /* [n] */
    int32_t *frame_flags;  /* [n] OPP_FLAG_* bits, may be NULL */
    /* Optional materialised up-sampled maps (DEVICE pointers), what the reference keeps in
     * upsample_conf/upsample_paf (src/paf.cpp:74-75) and the Python entry returns
     * (post_process.py:150).  NULL = not written (the skeletons are identical either way). */
    float *conf_up;        /* [n, 19, H, W] (CHW) or [n, H, W, 19] (up_layout HWC) */
    float *paf_up;         /* [n, 38, H, W] (CHW) or [n, H, W, 38] */
    int32_t up_layout;
    int32_t in_sync;       /* OPP_SYNC_*: ordering of device-resident conf / paf (and of conf_up / paf_up re-use) against
                            * the producer; ignored for host inputs */
    void *in_sync_obj;     /* cudaStream_t or cudaEvent_t, see OPP_SYNC_* */
} opp_batch_t;

typedef struct opp_handle_s *opp_handle_t;

/* Threading contract: a handle is used by ONE host thread at a time (like the reference's paf_processor instance,
 * whose scratch members make it non-reentrant: src/paf.cpp:74-77, examples/stream_detector.cpp:120-136); different
 * handles - on the same or on different GPUs - may be driven from different threads concurrently.  opp_host_alloc /
 * opp_host_free / opp_host_register / opp_host_unregister are thread-safe.  Every entry point leaves the caller's
 * current CUDA device unchanged. */

void opp_config_default(opp_config_t *cfg, int feat_h, int feat_w, int out_h, int out_w, int gauss_kernel_size);

/* replaces paf_processor_impl's constructor (src/paf.cpp:22-36) and peak_finder_t's (src/post-process.h:142-153) */
int opp_create(const opp_config_t *cfg, opp_handle_t *out);
void opp_destroy(opp_handle_t h);

/* replaces paf_processor_impl::operator() (src/paf.cpp:38-57) for n frames; returns when outputs are written */
int opp_process(opp_handle_t h, const opp_batch_t *batch);

/* Pipelined form: opp_submit enqueues a batch on the next free slot and returns a ticket;
 * opp_wait blocks until that batch's outputs are written.  Up to n_slots batches in flight. */
int opp_submit(opp_handle_t h, const opp_batch_t *batch, int *ticket);
int opp_wait(opp_handle_t h, int ticket);

/* Device time (ms, CUDA events on the slot's stream) of the last completed batch on a ticket's slot. */
float opp_last_batch_ms(opp_handle_t h, int ticket);
/* CUDA ordinal the handle lives on (resolves device = -1). */
int opp_device(opp_handle_t h);
/* Number of kernel launches issued by this handle so far. */
int64_t opp_launch_count(opp_handle_t h);
/* Memory-safety check (debug builds only).  A library built with -DOPP_DEBUG_BOUNDS (python -m openpose_plus_b200.build
 * --debug -> libopp_b200_dbg.so) checks every access to the arrays its kernels carve out of shared memory; this call waits
 * for the device, returns the first violation since the last call as out = {source line of the array in
 * csrc/opp_kernels.cu, index, size, number of violations} (all 0: none) and clears the record.  OPP_ERR_INVALID in a
 * release build. */
int opp_debug_bounds_report(opp_handle_t h, int32_t out[4]);
/* Test entry for the limb kernel's emulation of std::sort(greater on score) (src/paf.cpp:151-152; libstdc++'s introsort: with
 * tied scores its element movement decides the order): sorts cands[0..n) (host memory; cid1 / cid2 are carried along) in place
 * on the device.  mode 0 = the parallel form the kernel uses for up to 4096 candidates in shared memory, 1 = the sequential
 * emulation; threads = CTA size (multiple of 32, <= 256). */
int opp_debug_sort(opp_handle_t h, opp_conn_t *cands, int n, int mode, int threads);
/* Which peak kernel opp_create selected for this geometry and kernel size (a static string):
 *   "fast"        integer scale 8 or 4, Gaussian radius <= 2 x scale (k <= 33 at x8): reads the feature maps only
 *   "generic_rep" any other integer scale / larger kernels: replication-aware, reads the feature maps only
 *   "generic"     non-integer scales: 2-tap area-mode samples of the feature maps, staged per tile */
const char *opp_peak_kernel(opp_handle_t h);

/* Pinned host memory for opp_batch_t host buffers (pageable memory also works, slower). */
void *opp_host_alloc(size_t bytes);
void opp_host_free(void *p);
/* Flags for opp_host_alloc_ex.  WRITE_COMBINED memory is for INPUT buffers a producer only writes (the GPU reads it
 * over PCIe without snooping the CPU caches; CPU reads of it are very slow): never use it for result buffers. */
enum { OPP_HOST_DEFAULT = 0, OPP_HOST_WRITE_COMBINED = 1 };
void *opp_host_alloc_ex(size_t bytes, int flags);
/* Pins memory the caller already owns (e.g. a shared-memory segment several processes map, so that the ranks of a
 * multi-GPU job write their skeletons straight into one host buffer: the "host gather" of BASELINE.json configs[4]
 * without a copy).  p and bytes should be page-aligned.  Returns OPP_OK or OPP_ERR_CUDA. */
int opp_host_register(void *p, size_t bytes);
int opp_host_unregister(void *p);

/* Makes `stream` (a cudaStream_t of the caller) wait for the batch behind `ticket` without blocking the host:
 * the device-side counterpart of opp_wait for consumers of device-resident results (out_mem = OPP_MEM_DEVICE,
 * conf_up / paf_up).  The slot stays in flight until opp_wait(ticket) is called. */
int opp_stream_wait_ticket(opp_handle_t h, int ticket, void *stream);

/* Intermediates of the last batch processed on `ticket`'s slot, for parity tests.  Each call copies
 * up to cap elements of frame `frame` to host memory and returns the element count (<0 on error).
 *   OPP_DBG_PEAKS  -> opp_peak_t[]  (all_peaks in raster order, src/post-process.h:190-198)
 *   OPP_DBG_CONNS  -> opp_conn_t[]  of limb `index` in acceptance order (src/paf.cpp:154-173)
 *   OPP_DBG_PARTS  -> int32[18] part -> peak id of surviving human `index` (human_ref_t::parts) */
enum { OPP_DBG_PEAKS = 0, OPP_DBG_CONNS = 1, OPP_DBG_PARTS = 2, OPP_DBG_COUNTS = 3 };
int opp_debug_fetch(opp_handle_t h, int ticket, int what, int frame, int index, void *dst, int cap);

/* Stand-alone stages on device memory (used by tests and the bench to time kernels in isolation).
 * stream is a cudaStream_t (NULL = the handle's slot-0 stream). */
int opp_resize_device(opp_handle_t h, const float *src, int channels, int n_frames, float *dst, int dst_layout, void *stream);
/* Heat maps and PAFs of n frames in ONE launch (what opp_process issues when both outputs are requested). */
int opp_resize_pair_device(opp_handle_t h, const float *conf, const float *paf, int n_frames, float *conf_up, float *paf_up, int dst_layout,
                           void *stream);

/* Peak finding alone (smooth + NMS + raster-order peak list) on device feature maps, on the caller's
 * stream, writing the slot-0 scratch buffers: lets the bench time the kernel with its own events.
 * With conf_up/paf_up (and paf) non-NULL it runs the variant that also materialises the up-sampled
 * maps from inside the kernel, as opp_process does when both outputs are requested channels-first. */
int opp_peaks_device(opp_handle_t h, const float *conf, const float *paf, int n_frames, float *conf_up, float *paf_up, void *stream);

/* Device-side stopwatch over ALL slot streams: opp_timer_start records an event after everything
 * already enqueued; opp_timer_stop joins every slot stream, records, waits, and returns the elapsed
 * milliseconds between the two events (CUDA events, not wall clock). */
int opp_timer_start(opp_handle_t h);
float opp_timer_stop(opp_handle_t h);

/* Skeleton overlay of one human on an interleaved 8-bit image (1..4 channels, first three written as
 * R,G,B of the reference's palette): the role of draw_human (examples/vis.cpp:56-81) without OpenCV.
 * row_stride_bytes = 0 means width * channels.  Host code; needs no GPU and no handle. */
int opp_draw_human(uint8_t *image, int height, int width, int channels, ptrdiff_t row_stride_bytes, const opp_human_t *human, int thickness);

/* Calls opp_process(h, batch) `iters` times from C and writes each call's wall-clock duration in microseconds to
 * out_us[iters] (std::chrono::steady_clock around the call): the latency a C or C++ caller of this ABI sees, without
 * the interpreter overhead a Python loop adds to every call.  Returns the status of the first failing call. */
int opp_bench_latency(opp_handle_t h, const opp_batch_t *batch, int iters, float *out_us);

/* Host->device ceiling of this handle's input path, nothing else: `iters` times the same two cudaMemcpyAsync calls
 * opp_submit issues for a host batch of n_frames (conf, paf -> the slots' staging buffers, on the slots' streams), no
 * kernels; conf / paf hold n_batches consecutive batches that are cycled through; *out_ms = CUDA-event time over all
 * slot streams.  bench.py runs it on every rank at once to put a measured PCIe ceiling beside the end-to-end number. */
int opp_bench_h2d(opp_handle_t h, const float *conf, const float *paf, int n_frames, int n_batches, int iters, float *out_ms);

const char *opp_last_error(opp_handle_t h);
const char *opp_version(void);

/* The reference declares this and never defines it (include/openpose-plus.h:18-22).  Defined here on
 * top of the path above: runs one frame (feature maps of size height x width, up-sampled x8 like
 * every caller in the reference, gauss kernel 17) and prints each human like human_t::print. */
void process_conf_paf(int height, int width, int n_joins, int n_connections, const float *peaks_, const float *pafmap_);

#ifdef __cplusplus
}
#endif
#endif
