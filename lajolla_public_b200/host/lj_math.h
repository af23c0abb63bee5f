// lj_math.h -- double-precision vectors and 4x4 transforms of the host front end.  The reference parses and
// transforms everything in double (lajolla.h:23 Real = double; transform.cpp) and only the Embree buffers / our flat
// description are float, so the host side does the same arithmetic before rounding to fp32.
#pragma once
#include <cmath>

namespace ljhost {

constexpr double kPi = 3.14159265358979323846;

struct Vec2 { double x = 0, y = 0; };
struct Vec3 {
    double x = 0, y = 0, z = 0;
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(double s, Vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(Vec3 a, double s) { double inv = 1.0 / s; return {a.x * inv, a.y * inv, a.z * inv}; }  // vector.h:113-116
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) {  // vector.h:249-257: the zero vector stays zero
    double l = length(a);
    if (l <= 0) return {0, 0, 0};
    return a / l;
}
inline double radians(double deg) { return (kPi / 180.0) * deg; }
inline double degrees(double rad) { return (180.0 / kPi) * rad; }

struct Mat4 {
    double m[4][4];
    double &operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
    static Mat4 identity() {
        Mat4 r;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = i == j ? 1.0 : 0.0;
        return r;
    }
    static Mat4 zero() {
        Mat4 r;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = 0.0;
        return r;
    }
};
inline Mat4 operator*(const Mat4 &a, const Mat4 &b) {
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
// inverse by cofactors built from 2x2 sub-determinants; a singular matrix gives the zero matrix (matrix.h:207-209)
inline Mat4 inverse(const Mat4 &a) {
    const double (*m)[4] = a.m;
    double s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1], s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
    double s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3], s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
    double s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3], s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
    double c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    double c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    double c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    if (det == 0) return Mat4::zero();
    double id = 1.0 / det;
    Mat4 r;
    r.m[0][0] = (m[1][1] * c5 - m[1][2] * c4 + m[1][3] * c3) * id;
    r.m[0][1] = (-m[0][1] * c5 + m[0][2] * c4 - m[0][3] * c3) * id;
    r.m[0][2] = (m[3][1] * s5 - m[3][2] * s4 + m[3][3] * s3) * id;
    r.m[0][3] = (-m[2][1] * s5 + m[2][2] * s4 - m[2][3] * s3) * id;
    r.m[1][0] = (-m[1][0] * c5 + m[1][2] * c2 - m[1][3] * c1) * id;
    r.m[1][1] = (m[0][0] * c5 - m[0][2] * c2 + m[0][3] * c1) * id;
    r.m[1][2] = (-m[3][0] * s5 + m[3][2] * s2 - m[3][3] * s1) * id;
    r.m[1][3] = (m[2][0] * s5 - m[2][2] * s2 + m[2][3] * s1) * id;
    r.m[2][0] = (m[1][0] * c4 - m[1][1] * c2 + m[1][3] * c0) * id;
    r.m[2][1] = (-m[0][0] * c4 + m[0][1] * c2 - m[0][3] * c0) * id;
    r.m[2][2] = (m[3][0] * s4 - m[3][1] * s2 + m[3][3] * s0) * id;
    r.m[2][3] = (-m[2][0] * s4 + m[2][1] * s2 - m[2][3] * s0) * id;
    r.m[3][0] = (-m[1][0] * c3 + m[1][1] * c1 - m[1][2] * c0) * id;
    r.m[3][1] = (m[0][0] * c3 - m[0][1] * c1 + m[0][2] * c0) * id;
    r.m[3][2] = (-m[3][0] * s3 + m[3][1] * s1 - m[3][2] * s0) * id;
    r.m[3][3] = (m[2][0] * s3 - m[2][1] * s1 + m[2][2] * s0) * id;
    return r;
}

// transform.cpp:5-88 (pbrt-style constructors; look_at builds a camera-to-world frame with +z = viewing direction)
inline Mat4 translate(Vec3 d) { Mat4 r = Mat4::identity(); r(0, 3) = d.x; r(1, 3) = d.y; r(2, 3) = d.z; return r; }
inline Mat4 scale(Vec3 s) { Mat4 r = Mat4::identity(); r(0, 0) = s.x; r(1, 1) = s.y; r(2, 2) = s.z; return r; }
inline Mat4 rotate(double angle_deg, Vec3 axis) {
    Vec3 a = normalize(axis);
    double s = std::sin(radians(angle_deg)), c = std::cos(radians(angle_deg));
    Mat4 m = Mat4::identity();
    m(0, 0) = a.x * a.x + (1 - a.x * a.x) * c; m(0, 1) = a.x * a.y * (1 - c) - a.z * s; m(0, 2) = a.x * a.z * (1 - c) + a.y * s;
    m(1, 0) = a.x * a.y * (1 - c) + a.z * s; m(1, 1) = a.y * a.y + (1 - a.y * a.y) * c; m(1, 2) = a.y * a.z * (1 - c) - a.x * s;
    m(2, 0) = a.x * a.z * (1 - c) - a.y * s; m(2, 1) = a.y * a.z * (1 - c) + a.x * s; m(2, 2) = a.z * a.z + (1 - a.z * a.z) * c;
    return m;
}
inline Mat4 look_at(Vec3 pos, Vec3 look, Vec3 up) {
    Vec3 dir = normalize(look - pos);
    Vec3 left = normalize(cross(normalize(up), dir));
    Vec3 new_up = cross(dir, left);
    Mat4 m = Mat4::identity();
    m(0, 0) = left.x; m(1, 0) = left.y; m(2, 0) = left.z;
    m(0, 1) = new_up.x; m(1, 1) = new_up.y; m(2, 1) = new_up.z;
    m(0, 2) = dir.x; m(1, 2) = dir.y; m(2, 2) = dir.z;
    m(0, 3) = pos.x; m(1, 3) = pos.y; m(2, 3) = pos.z;
    return m;
}
inline Mat4 perspective(double fov_deg) {
    double cot = 1.0 / std::tan(radians(fov_deg / 2.0));
    Mat4 m = Mat4::zero();
    m(0, 0) = cot; m(1, 1) = cot; m(2, 2) = 1; m(2, 3) = -1; m(3, 2) = 1;
    return m;
}
inline Vec3 xform_point(const Mat4 &t, Vec3 p) {
    double x = t(0, 0) * p.x + t(0, 1) * p.y + t(0, 2) * p.z + t(0, 3);
    double y = t(1, 0) * p.x + t(1, 1) * p.y + t(1, 2) * p.z + t(1, 3);
    double z = t(2, 0) * p.x + t(2, 1) * p.y + t(2, 2) * p.z + t(2, 3);
    double w = t(3, 0) * p.x + t(3, 1) * p.y + t(3, 2) * p.z + t(3, 3);
    double inv_w = 1.0 / w;
    return {x * inv_w, y * inv_w, z * inv_w};
}
inline Vec3 xform_vector(const Mat4 &t, Vec3 v) {
    return {t(0, 0) * v.x + t(0, 1) * v.y + t(0, 2) * v.z, t(1, 0) * v.x + t(1, 1) * v.y + t(1, 2) * v.z,
            t(2, 0) * v.x + t(2, 1) * v.y + t(2, 2) * v.z};
}
// normals go through the transpose of the inverse, then are renormalised (transform.cpp:95-100)
inline Vec3 xform_normal(const Mat4 &inv, Vec3 n) {
    return normalize(Vec3{inv(0, 0) * n.x + inv(1, 0) * n.y + inv(2, 0) * n.z, inv(0, 1) * n.x + inv(1, 1) * n.y + inv(2, 1) * n.z,
                          inv(0, 2) * n.x + inv(1, 2) * n.y + inv(2, 2) * n.z});
}

}  // namespace ljhost
