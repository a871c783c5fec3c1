// draw_human on the new structs (the reference's examples/vis.h:7, examples/vis.cpp:56-81).  The reference
// draws into a cv::Mat; this header draws into any interleaved 8-bit image through the C-ABI
// (opp_draw_human, include/opp_b200.h) and adds the cv::Mat overload when OpenCV's core header was
// included first.
#pragma once
#include <cstddef>
#include <cstdint>

#include <opp_b200.h>
#include <openpose-plus/human.h>

inline void draw_human(uint8_t *image, int height, int width, int channels, const human_t &human, int thickness = 2)
{
    // human_t and opp_human_t share one 292-byte layout (static_assert in csrc/paf_processor.cpp)
    opp_draw_human(image, height, width, channels, 0, reinterpret_cast<const opp_human_t *>(&human), thickness);
}

#ifdef OPENCV_CORE_HPP
inline void draw_human(cv::Mat &img, const human_t &human)
{
    opp_draw_human(img.data, img.rows, img.cols, img.channels(), (ptrdiff_t)img.step, reinterpret_cast<const opp_human_t *>(&human), 2);
}
#endif
