// Post-processing stage of a batch detector written against the reference's public API
// (what examples/pose_detector.cpp:96-116 does after TensorRT has produced the feature maps), plus the
// batch extension.  Synthetic all-zero maps with one blob stand in for the CNN output.
//   g++ -std=c++14 -Iinclude examples/batch_detector.cpp -Lopenpose_plus_b200 -l:libopp_b200.so \
//       -Wl,-rpath,$PWD/openpose_plus_b200 -o batch_detector
#include <cstdio>
#include <memory>
#include <vector>

#include <openpose-plus.h>

int main()
{
    const int fh = 46, fw = 54, H = 368, W = 432, batch = 8;
    std::vector<float> confs((size_t)batch * n_joins * fh * fw, 0.f), pafs((size_t)batch * 2 * n_connections * fh * fw, 0.f);
    // frame-by-frame, exactly like the reference
    std::unique_ptr<paf_processor> process_paf(create_paf_processor(fh, fw, H, W, n_joins, n_connections, 17));
    for (int i = 0; i < batch; ++i) {
        const auto humans = (*process_paf)(confs.data() + (size_t)i * n_joins * fh * fw, pafs.data() + (size_t)i * 2 * n_connections * fh * fw, true);
        std::printf("frame %d: %zu humans\n", i, humans.size());
        for (const auto &h : humans) h.print();
    }
    // the whole batch in one call
    std::unique_ptr<paf_batch_processor> batch_paf(create_paf_batch_processor(fh, fw, H, W, 17, batch, -1));
    const auto all = batch_paf->process_batch(confs.data(), pafs.data(), batch, /*device_memory=*/false);
    std::printf("batch call: %zu frames\n", all.size());
    return 0;
}
