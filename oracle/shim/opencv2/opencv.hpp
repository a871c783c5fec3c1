// TEST INFRASTRUCTURE (oracle). Minimal stand-in for the OpenCV surface used by the reference at
// src/post-process.h:39-47,64-70 (cv::Size, cv::Mat over caller memory, cv::DataType, CV_INTER_AREA,
// cv::resize, cv::GaussianBlur).  OpenCV C++ headers/libs are not installed in this image; the two
// functions are defined in oracle/cv_standin.cpp as scalar restatements validated against cv2 4.13
// (IPP off, setUseOptimized(False)).
#pragma once
#include <cstddef>

#define CV_32F 5
#define CV_INTER_AREA 3

namespace cv
{
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

template <typename T> struct DataType;
template <> struct DataType<float> {
    static const int type = CV_32F;
};

class Mat
{
  public:
    Mat(Size s, int type, void *data) : size_(s), type_(type), data_(data) {}
    Size size() const { return size_; }
    int type() const { return type_; }
    void *ptr() const { return data_; }

  private:
    Size size_;
    int type_;
    void *data_;
};

void resize(const Mat &src, Mat &dst, Size dsize, double fx, double fy,
            int interpolation);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX);
}  // namespace cv
