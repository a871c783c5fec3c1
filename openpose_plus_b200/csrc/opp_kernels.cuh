// Kernel parameter blocks and launchers of the openpose-plus post-processing path on sm_100a.
// Reference file:line citations are relative to /root/reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/opp_b200.h"

#define OPP_MAX_KSIZE 63
#define OPP_THREADS 256
#define K2_FAST_MAX_GROUPS 7 // column groups (warps) per CTA of the integer-scale peak kernel: 7 x 62 decided columns >= 432

// Geometry shared by every stage.  S > 0 means both axes scale by the same integer factor
// (every configuration in BASELINE.json is x8): INTER_AREA up-sampling is then exact pixel
// replication and the fast kernels index the feature maps directly.  Otherwise the area-mode
// coefficient tables (cv::resize, see oracle/opp_oracle.c orc_resize_coeffs) drive a 2x2-tap sample.
struct OppGeom {
    int h, w, H, W;
    int S;
    int K, R;
    int xmax;           // first output column whose right tap leaves the source (single tap from there)
    const int *xofs;    // [W]
    const float *alpha; // [W][2]
    const int *yofs;    // [H]
    const float *beta;  // [H][2]
};

// frame-level counters, cleared by one memset per batch
struct OppCounters {
    int *pk_cnt;  // [n][18] peaks appended per (frame, part)
    int *k2_done; // [n]     unused since the peak kernels have no per-frame epilogue (kept for the counter layout)
    int *k3_done; // [n]     limbs of the frame that finished matching
};

struct K2Params {
    OppGeom g;
    const float *conf;    // [n,19,h,w] feature maps
    const float *conf_up; // [n,19,H,W] materialised maps (INPUT of the generic kernel only)
    // fused materialisation (fast kernel): when up_conf != nullptr the kernel also writes both up-sampled tensors
    const float *paf;     // [n,38,h,w]
    float *up_conf;       // [n,19,H,W]
    float *up_paf;        // [n,38,H,W]
    int nxs, nys, tw, th; // tiling in feature-map units (fast kernel) / output tiles (generic)
    int capP;
    OppCounters cnt;
    int *pk_key;          // [n][18][capP] y*W+x, unordered
    opp_peak_t *peaks;    // [n][18*capP] raster order
    int *part_ofs;        // [n][19]
    int *flags;           // [n]
    float taps[OPP_MAX_KSIZE + 1];
    float thresh;
    unsigned long long *times; // optional [ctas][8] %globaltimer phase stamps (debug, single frame), may be null
    int border_zero;   // taps outside the image read 0 (Python-path variant) instead of REFLECT_101
    float skip_thresh; // blocks whose 3x3 feature neighbourhood stays <= this cannot hold a peak; -inf disables the skip
};

struct K3Params {
    OppGeom g;
    const float *paf; // [n,38,h,w]
    // peaks: the peak kernel's unordered keys in, the reference's all_peaks (raster order, ids) out
    const int *pk_key;      // [n][18][capP] y*W+x, unordered
    const float *conf;      // [n,19,h,w] feature maps (peak scores)
    const float *conf_up;   // [n,19,H,W] materialised map, only at non-integer scales
    opp_peak_t *peaks;      // [n][18*capP] written by this kernel
    int *part_ofs;          // [n][19]      written by this kernel
    int capP, capC, capH;
    OppCounters cnt;
    float *cand_scratch;   // [n][19][2][capC][3] when candidates do not fit shared memory
    opp_conn_t *conns;     // [n][19][capP]
    int *n_conns;          // [n][19]
    opp_human_t *humans;   // [n][capH]
    int *n_humans;         // [n]
    int *flags;            // [n] device flag words, OR-ed by every stage
    int *flags_out;        // [n] optional: final flag word of the frame, written once by the assembly (may be mapped host memory)
    int *href_parts;       // [n][capH][18]
    int *stats;            // [n][4] partial humans, merges, total candidates, pairs that passed the quick PAF test
    int paf_in_smem, cand_in_smem, score_in_smem, conns_in_smem, owner_in_smem;
    // shared-memory carve-up (byte offsets)
    int off_paf, off_pk, off_cand, off_used, off_misc, off_keys, off_href, off_score, off_conn, off_keep, off_owner;
    int off_cand1;          // second candidate buffer (shares its bytes with the PAF tile, which is dead by then)
    int conn_cap, pk_cap;   // entries of the assembly's staging areas for connections / peaks (frames beyond them: slower forms)
    int off_steps, steps_in_smem; // [max(H, W)] floats: d / 10.f, the PAF sampling steps as a table (built by every limb CTA)
    int off_weak, weak_in_smem;   // [ceil(h w / 32)] words: feature cells whose PAF cannot lift a sample over THRESH_VECTOR_SCORE
    int cand_unordered;           // candidates are appended without block-wide ordering (capC <= 4096, shared memory)
    int off_surv, surv_cap; // [surv_cap] ints: pairs that passed the quick PAF test (then, as bytes, the matching's per-candidate state)
    float thr_vec, thr_human;
    // completion word for the latency path: every frame's assembly bumps batch_done; the one that completes the batch
    // writes done_tag to host_done (mapped pinned memory) after a system-scope fence, so the host can spin on it
    // instead of waiting for the kernel to retire and an event to be signalled.  host_done == nullptr: unused.
    int *batch_done;
    int *host_done;
    int done_tag, done_frames;
    int paf_early;             // paf is mapped pinned host memory: each CTA requests its tile before griddepcontrol.wait
    int true_index;            // assembly indexes humans by position (pafprocess) instead of by stored id (src/paf.cpp:198,204)
    unsigned long long *times; // optional [n][19][12] %globaltimer stamps of the phases (debug), may be null
};

struct K1Params {
    OppGeom g;
    const float *src; // [n,C,h,w]
    float *dst;       // [n,C,H,W] or [n,H,W,C]
    int C, n;
    int layout;
    // optional second tensor handled by the same launch (heat maps + PAFs in one grid); C2 == 0: none
    const float *src2;
    float *dst2;
    int C2;
};

size_t k2_fast_smem_bytes(const OppGeom &g, int tw, int th);
bool k2_fast_supported(const OppGeom &g, bool border_zero);
cudaError_t launch_k2_fast(const K2Params &p, int n_frames, cudaStream_t st, bool pdl = false);
cudaError_t launch_k2_generic(const K2Params &p, int n_frames, cudaStream_t st);
cudaError_t launch_k2_generic_rep(const K2Params &p, int n_frames, cudaStream_t st); // integer scale: reads the feature maps
cudaError_t launch_k3(const K3Params &p, int n_frames, size_t smem, cudaStream_t st, bool pdl = false);
cudaError_t launch_k1(const K1Params &p, cudaStream_t st);
cudaError_t launch_hwc_to_chw(const float *src, float *dst, int n, int C, int h, int w, cudaStream_t st);
cudaError_t launch_ingest(const float *src0, float *dst0, size_t n0, const float *src1, float *dst1, size_t n1, int *counters, int n_counters,
                          cudaStream_t st);
cudaError_t opp_kernels_init(int max_smem_optin);
cudaError_t opp_kernels_bounds_report(int out[4], bool reset);
cudaError_t launch_debug_sort(void *cands, int n, int mode, int threads, cudaStream_t st); // cands: n x {int32 i1, int32 i2, float score}
