// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE (see oracle/__init__.py).
// Just enough of the OpenCV C++ surface for the reference's RetinaFace.cpp to COMPILE
// unmodified (OpenCV C++ headers are absent from this image).  Only cv::Vec4f is ever
// executed by oracle/_ref; every image routine aborts if reached, because oracle/_ref
// drives the reference's post-process (postProcess / nms / anchors), not its preprocess.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
// the real opencv2/opencv.hpp pulls these in transitively; the reference relies on that
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
namespace cv {
[[noreturn]] inline void shim_unreachable(const char *what) {
    std::fprintf(stderr, "oracle/shim: cv::%s is a compile-only stub\n", what);
    std::abort();
}
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Vec4f {
    float val[4];
    Vec4f() : val{0, 0, 0, 0} {}
    Vec4f(float a, float b, float c, float d) : val{a, b, c, d} {}
    float &operator[](int i) { return val[i]; }
    const float &operator[](int i) const { return val[i]; }
};
struct Mat {
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    Mat() {}
    Mat(int r, int c, int /*type*/, void *d) : rows(r), cols(c), data((unsigned char *)d) {}
    bool empty() const { return data == nullptr; }
    Mat clone() const { return *this; }
    void convertTo(Mat &, int) const { shim_unreachable("Mat::convertTo"); }
};
enum { BORDER_CONSTANT = 0 };
inline void resize(const Mat &, Mat &, Size, double = 0, double = 0) { shim_unreachable("resize"); }
inline void copyMakeBorder(const Mat &, Mat &, int, int, int, int, int, const Scalar & = Scalar()) { shim_unreachable("copyMakeBorder"); }
inline void cvtColor(const Mat &, Mat &, int) { shim_unreachable("cvtColor"); }
inline void split(const Mat &, std::vector<Mat> &) { shim_unreachable("split"); }
inline int64_t getTickCount() { return 0; }
inline double getTickFrequency() { return 1.0; }
}  // namespace cv
#define CV_32FC1 5
#define CV_32FC3 21
#define CV_BGR2RGB 4
