// C++ public API of openpose-plus, B200-native build.  Source-compatible with the reference's
// include/openpose-plus.hpp:42-64 for the post-processing path: callers keep
//     std::unique_ptr<paf_processor> p(create_paf_processor(fh, fw, H, W, 19, 19, ksize));
//     std::vector<human_t> humans = (*p)(heatmap, paf, use_gpu);
// The CNN runner (pose_detection_runner, TensorRT; reference include/openpose-plus.hpp:9-36) is out of scope of this
// build: it is DECLARED below, unchanged, so that translation units of the reference that implement or call it
// (src/uff-runner.cpp, examples/*.cpp) keep compiling against this header; libopp_b200.so does not define it.
#pragma once
#include <string>
#include <vector>

#include <openpose-plus/human.h>

// Feature-map producer: image batch in, heat maps [batch, 19, H', W'] and PAFs [batch, 38, H', W'] out.
class pose_detection_runner
{
  public:
    // inputs: one pointer to float[max_batch_size * 3 * H * W]; outputs: two pointers (heat maps, PAFs)
    virtual void operator()(const std::vector<void *> &inputs, const std::vector<void *> &outputs, int batchSize = 1) = 0;

    virtual ~pose_detection_runner() {}
};

// Defined by the reference's TensorRT runner (src/uff-runner.cpp), not by this library.
pose_detection_runner *create_pose_detection_runner(const std::string &model_file, int input_height, int input_width,
                                                    int max_batch_size, bool use_f16);

class paf_processor
{
  public:
    // heatmap: host float[19, feat_h, feat_w]; paf: host float[38, feat_h, feat_w].
    // use_gpu is accepted and ignored: this build always runs on the GPU (no CPU fallback).
    virtual std::vector<human_t> operator()(const float *heatmap, const float *paf, bool use_gpu) = 0;

    virtual ~paf_processor() {}
};

// input_height/input_width: feature-map size; height/width: size the maps are up-sampled to
// (normally the image size); n_joins and n_connections must be 19.
paf_processor *create_paf_processor(int input_height, int input_width, int height, int width, int n_joins,
                                    int n_connections, int gauss_kernel_size);

// Batch extension (not in the reference): the same object also processes many frames per call.
class paf_batch_processor : public paf_processor
{
  public:
    // confs: [n, 19, fh, fw], pafs: [n, 38, fh, fw]; device_memory says where they live.
    virtual std::vector<std::vector<human_t>> process_batch(const float *confs, const float *pafs, int n_frames,
                                                            bool device_memory) = 0;
};
paf_batch_processor *create_paf_batch_processor(int input_height, int input_width, int height, int width,
                                                int gauss_kernel_size, int max_batch, int device);
