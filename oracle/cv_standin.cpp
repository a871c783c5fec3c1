// TEST INFRASTRUCTURE (oracle).  Definitions of the two OpenCV functions the reference's
// src/post-process.h calls (cv::resize at :46-47, cv::GaussianBlur at :69-70), linked into
// oracle/_ref/libopp_ref.so next to the reference's unmodified src/paf.cpp.  Both forward to the
// scalar restatements in opp_oracle.c, which tests/test_oracle_golden.py (test_against_live_cv2 and the committed tests/golden/opencv_pieces.npz) holds bit-identical to
// cv2 4.13 (IPP off, setUseOptimized(False)).
#include <cstdio>
#include <cstdlib>

#include <opencv2/opencv.hpp>

#include "opp_oracle.h"

namespace cv
{
void resize(const Mat &src, Mat &dst, Size dsize, double, double, int interpolation)
{
    if (interpolation != CV_INTER_AREA || src.type() != CV_32F) {
        std::fprintf(stderr, "cv_standin: only INTER_AREA on CV_32F is restated\n");
        std::abort();
    }
    if (orc_resize_area_up((const float *)src.ptr(), src.size().height, src.size().width,
                           (float *)dst.ptr(), dsize.height, dsize.width)) {
        std::fprintf(stderr, "cv_standin: INTER_AREA down-sampling is not restated\n");
        std::abort();
    }
}

void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX)
{
    if (ksize.width != ksize.height ||
        orc_gauss_blur((const float *)src.ptr(), src.size().height, src.size().width, ksize.width,
                       sigmaX, (float *)dst.ptr())) {
        std::fprintf(stderr, "cv_standin: unsupported Gaussian kernel\n");
        std::abort();
    }
}
}  // namespace cv
