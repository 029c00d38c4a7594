// Stand-ins for the Eigen types that appear in the PPF API when Eigen itself is not installed:
// Matrix4f (column-major storage like Eigen), Vector3f, Affine3f.  Not a linear-algebra library.
#pragma once

#include <cstring>

namespace Eigen {

struct Vector3f {
    float v[3] = {0, 0, 0};
    Vector3f() = default;
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
};

struct Vector4f {
    float v[4] = {0, 0, 0, 0};
    Vector4f() = default;
    Vector4f(float x, float y, float z, float w) : v{x, y, z, w} {}
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};

struct Matrix4f {
    float m[16];  // column-major
    Matrix4f() { std::memset(m, 0, sizeof(m)); }
    static Matrix4f Identity() {
        Matrix4f I;
        I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1.0f;
        return I;
    }
    static Matrix4f fromRowMajor(const float *r) {
        Matrix4f M;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) M.m[j * 4 + i] = r[i * 4 + j];
        return M;
    }
    float &operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    float *data() { return m; }
    const float *data() const { return m; }
    bool operator==(const Matrix4f &o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
    bool operator!=(const Matrix4f &o) const { return !(*this == o); }
    bool isIdentity() const { return *this == Identity(); }
};

struct Affine3f {
    Matrix4f M = Matrix4f::Identity();
    Affine3f() = default;
    explicit Affine3f(const Matrix4f &m) : M(m) {}
    const Matrix4f &matrix() const { return M; }
    Matrix4f &matrix() { return M; }
    Vector3f translation() const { return Vector3f(M(0, 3), M(1, 3), M(2, 3)); }
    float operator()(int r, int c) const { return M(r, c); }
};

}  // namespace Eigen
