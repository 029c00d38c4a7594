// Stand-ins for the few OpenCV core types the surface_matching API exposes (cv::Mat of CV_32F rows, cv::Matx44d,
// cv::Vec3d, cv::Ptr) for builds without OpenCV; with OpenCV on the include path its own types are used instead.
// Not an image library: a row-major float matrix with shared storage, and a 4x4 double matrix.
#pragma once

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define B200PPF_HAVE_OPENCV 1
#endif
#endif

#ifndef B200PPF_HAVE_OPENCV
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_32F
#define CV_32F 5
#define CV_32FC1 5
#endif

namespace cv {

template <typename T>
using Ptr = std::shared_ptr<T>;

struct Vec3d {
    double val[3] = {0, 0, 0};
    Vec3d() = default;
    Vec3d(double x, double y, double z) : val{x, y, z} {}
    double &operator[](int i) { return val[i]; }
    double operator[](int i) const { return val[i]; }
};

struct Matx44d {
    double val[16];  // row-major, like cv::Matx
    Matx44d() { std::memset(val, 0, sizeof(val)); }
    explicit Matx44d(const double *v) { std::memcpy(val, v, sizeof(val)); }
    static Matx44d eye() {
        Matx44d m;
        m.val[0] = m.val[5] = m.val[10] = m.val[15] = 1.0;
        return m;
    }
    double &operator()(int r, int c) { return val[4 * r + c]; }
    double operator()(int r, int c) const { return val[4 * r + c]; }
    Matx44d operator*(const Matx44d &o) const {
        Matx44d m;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double s = 0.0;
                for (int k = 0; k < 4; ++k) s += val[4 * r + k] * o.val[4 * k + c];
                m.val[4 * r + c] = s;
            }
        return m;
    }
};

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int rows_, int cols_, int type_) : rows(rows_), cols(cols_), type_(type_), data_(std::make_shared<std::vector<float>>((std::size_t)rows_ * cols_)) {}
    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    template <typename T>
    T *ptr(int i = 0) { return reinterpret_cast<T *>(data_->data() + (std::size_t)i * cols); }
    template <typename T>
    const T *ptr(int i = 0) const { return reinterpret_cast<const T *>(data_->data() + (std::size_t)i * cols); }
    template <typename T>
    T &at(int i, int j) { return ptr<T>(i)[j]; }
    Mat clone() const {
        Mat m(rows, cols, type_);
        if (data_) *m.data_ = *data_;
        return m;
    }

private:
    int type_ = CV_32F;
    std::shared_ptr<std::vector<float>> data_;
};

}  // namespace cv
#endif  // !B200PPF_HAVE_OPENCV
