// A caller written against the reference's public API only (create_paf_processor / paf_processor /
// human_t, include/openpose-plus.hpp) -- what examples/pose_detector.cpp:67-69,99-102 does per frame.
//   dropin_main <in.bin> <out.bin>
// in.bin : int32 geom[5] = {feat_h, feat_w, out_h, out_w, ksize}, int32 n_frames, then per frame
//          float conf[19*h*w], float paf[38*h*w]
// out.bin: per frame int32 n_humans, then n_humans * sizeof(human_t) bytes
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <openpose-plus.h>

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *fi = std::fopen(argv[1], "rb"), *fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) return 3;
    int geom[5], n = 0;
    if (std::fread(geom, sizeof(int), 5, fi) != 5 || std::fread(&n, sizeof(int), 1, fi) != 1) return 4;
    std::unique_ptr<paf_processor> process_paf(create_paf_processor(geom[0], geom[1], geom[2], geom[3], n_joins, n_connections, geom[4]));
    const size_t nc = (size_t)n_joins * geom[0] * geom[1], np = (size_t)2 * n_connections * geom[0] * geom[1];
    std::vector<float> conf(nc), paf(np);
    for (int f = 0; f < n; ++f) {
        if (std::fread(conf.data(), sizeof(float), nc, fi) != nc || std::fread(paf.data(), sizeof(float), np, fi) != np) return 5;
        const std::vector<human_t> humans = (*process_paf)(conf.data(), paf.data(), /*use_gpu=*/true);
        const int m = (int)humans.size();
        std::fwrite(&m, sizeof(int), 1, fo);
        if (m) std::fwrite(humans.data(), sizeof(human_t), m, fo);
    }
    // the batch extension on the same object type
    std::fclose(fi);
    std::fclose(fo);
    return 0;
}
