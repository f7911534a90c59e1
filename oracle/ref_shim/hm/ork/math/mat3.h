/* Shim standing in for Ork's ork/math/mat3.h (row-major 3x3), see vec3.h.  Test infrastructure. */
#ifndef ORC_SHIM_ORK_MAT3_H
#define ORC_SHIM_ORK_MAT3_H
#include "ork/math/vec3.h"
namespace ork {
template <typename T> struct mat3 {
    T m[3][3];
    mat3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = 0; }
    mat3(T a, T b, T c, T d, T e, T f, T g, T h, T i)
    { m[0][0] = a; m[0][1] = b; m[0][2] = c; m[1][0] = d; m[1][1] = e; m[1][2] = f; m[2][0] = g; m[2][1] = h; m[2][2] = i; }
    vec3<T> operator*(const vec3<T> &v) const
    { return vec3<T>(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                     m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z); }
    mat3 operator*(const mat3 &o) const
    { mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j]; return r; }
    mat3 transpose() const { return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]); }
};
typedef mat3<float> mat3f;
typedef mat3<double> mat3d;
}
using namespace ork;
#endif
