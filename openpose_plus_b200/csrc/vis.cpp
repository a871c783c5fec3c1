// Skeleton overlay on a raw interleaved 8-bit image: the role of the reference's draw_human
// (/root/reference examples/vis.cpp:56-81, colours :30-50) without OpenCV -- a thick line per limb of
// COCOPAIRS whose two parts are present, then a dot per present part.  Host code, outside the hot path.
// Rasterisation is a plain distance test (pixel centre within thickness/2 + 0.5 of the segment); it is
// not claimed to be pixel-identical to cv::line / cv::circle.
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "../../include/opp_b200.h"

namespace
{
const uint8_t kColors[19][3] = {{255, 0, 0},   {255, 85, 0},  {255, 170, 0}, {255, 255, 0}, {170, 255, 0}, {85, 255, 0},  {0, 255, 0},
                                {0, 255, 85},  {0, 255, 170}, {0, 255, 255}, {0, 170, 255}, {0, 85, 255},  {0, 0, 255},   {85, 0, 255},
                                {170, 0, 255}, {255, 0, 255}, {255, 0, 170}, {255, 0, 85},  {127, 127, 127}};
// include/openpose-plus/coco.h:11-32
const int kPairs[OPP_N_PAIRS][2] = {{1, 2},   {1, 5},   {2, 3}, {3, 4},  {5, 6},   {6, 7},  {1, 8},   {8, 9},  {9, 10}, {1, 11},
                                    {11, 12}, {12, 13}, {1, 0}, {0, 14}, {14, 16}, {0, 15}, {15, 17}, {2, 16}, {5, 17}};

struct View {
    uint8_t *px;
    int H, W, C;
    ptrdiff_t stride;
    void put(int x, int y, const uint8_t *rgb) const
    {
        uint8_t *q = px + y * stride + (ptrdiff_t)x * C;
        for (int c = 0; c < C && c < 3; ++c) q[c] = rgb[c];
    }
};

// every pixel whose centre lies within `rad` of the segment (x0,y0)-(x1,y1)
void capsule(const View &v, float x0, float y0, float x1, float y1, float rad, const uint8_t *rgb)
{
    const int xa = std::max(0, (int)std::floor(std::min(x0, x1) - rad)), xb = std::min(v.W - 1, (int)std::ceil(std::max(x0, x1) + rad));
    const int ya = std::max(0, (int)std::floor(std::min(y0, y1) - rad)), yb = std::min(v.H - 1, (int)std::ceil(std::max(y0, y1) + rad));
    const float dx = x1 - x0, dy = y1 - y0, len2 = dx * dx + dy * dy;
    for (int y = ya; y <= yb; ++y)
        for (int x = xa; x <= xb; ++x) {
            float t = len2 > 0.f ? ((x - x0) * dx + (y - y0) * dy) / len2 : 0.f;
            t = std::min(1.f, std::max(0.f, t));
            const float ex = x - (x0 + t * dx), ey = y - (y0 + t * dy);
            if (ex * ex + ey * ey <= rad * rad) v.put(x, y, rgb);
        }
}
}  // namespace

extern "C" int opp_draw_human(uint8_t *image, int height, int width, int channels, ptrdiff_t row_stride_bytes, const opp_human_t *human,
                              int thickness)
{
    if (!image || !human || height <= 0 || width <= 0 || channels < 1 || channels > 4 || thickness < 1) return OPP_ERR_INVALID;
    const View v{image, height, width, channels, row_stride_bytes ? row_stride_bytes : (ptrdiff_t)width * channels};
    for (int pair_id = 0; pair_id < OPP_N_PAIRS; ++pair_id) {
        const opp_body_part_t &a = human->parts[kPairs[pair_id][0]], &b = human->parts[kPairs[pair_id][1]];
        // cv::Point(p.x, p.y) truncates the float coordinates (examples/vis.cpp:68)
        if (a.has_value && b.has_value) capsule(v, (float)(int)a.x, (float)(int)a.y, (float)(int)b.x, (float)(int)b.y, 0.5f * thickness, kColors[pair_id]);
    }
    for (int part = 0; part < OPP_N_PARTS; ++part) {
        const opp_body_part_t &p = human->parts[part];
        // cv::circle(img, centre, radius = thickness, colour, line width = thickness): a ring of mean radius
        // `thickness` and width `thickness`, i.e. a disc of radius 1.5 * thickness with a pin-hole at most
        if (p.has_value) capsule(v, (float)(int)p.x, (float)(int)p.y, (float)(int)p.x, (float)(int)p.y, 1.5f * thickness, kColors[part]);
    }
    return OPP_OK;
}
