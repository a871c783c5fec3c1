// Result types of the openpose-plus public API (drop-in for the reference's
// include/openpose-plus/human.h:8-34).  Layout is ABI: sizeof(body_part_t) == 16,
// sizeof(human_t) == 292, identical to opp_body_part_t / opp_human_t in opp_b200.h.
#pragma once
#include <cstdio>
#include <vector>

#include <openpose-plus/coco.h>

struct body_part_t {
    bool has_value;
    float x;
    float y;
    float score;

    body_part_t() : has_value(false), x(0), y(0), score(0) {}
};

template <int J> struct human_t_ {
    body_part_t parts[J];
    float score;

    human_t_() : score(0) {}

    void print() const
    {
        for (int i = 0; i < J; ++i)
            if (parts[i].has_value)
                std::printf("BodyPart:%d-(%.2f, %.2f) score=%.2f ", i, parts[i].x, parts[i].y, parts[i].score);
        std::printf("score=%.2f\n", score);
    }
};

using human_t = human_t_<COCO_N_PARTS>;
