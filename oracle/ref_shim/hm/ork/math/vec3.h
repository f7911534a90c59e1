/* Shim standing in for Ork's ork/math/vec3.h, so that preprocess/terrain/ApertureMipmap.{h,cpp} (linked in because
 * Preprocess.cpp refers to it; the aperture builder itself is outside the path and never called) compile unchanged.
 * Test infrastructure. */
#ifndef ORC_SHIM_ORK_VEC3_H
#define ORC_SHIM_ORK_VEC3_H
#include "ork/core/Object.h"
namespace ork {
template <typename T> struct vec3 {
    T x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(T x, T y, T z) : x(x), y(y), z(z) {}
    template <typename U> vec3(const vec3<U> &v) : x((T) v.x), y((T) v.y), z((T) v.z) {}
    vec3 operator+(const vec3 &v) const { return vec3(x + v.x, y + v.y, z + v.z); }
    vec3 operator-(const vec3 &v) const { return vec3(x - v.x, y - v.y, z - v.z); }
    vec3 operator-() const { return vec3(-x, -y, -z); }
    vec3 operator*(T s) const { return vec3(x * s, y * s, z * s); }
    vec3 operator/(T s) const { return vec3(x / s, y / s, z / s); }
    vec3 &operator+=(const vec3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
    T dotproduct(const vec3 &v) const { return x * v.x + y * v.y + z * v.z; }
    T length() const { return (T) sqrt((double) (x * x + y * y + z * z)); }
    T squaredLength() const { return x * x + y * y + z * z; }
    vec3 normalize() const { T l = length(); return vec3(x / l, y / l, z / l); }
    vec3 normalize(T n) const { T l = length() / n; return vec3(x / l, y / l, z / l); }
    vec3 crossProduct(const vec3 &v) const { return vec3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }
    template <typename U> vec3<U> cast() const { return vec3<U>((U) x, (U) y, (U) z); }
    static const vec3 ZERO, UNIT_X, UNIT_Y, UNIT_Z;
};
template <typename T> const vec3<T> vec3<T>::ZERO(0, 0, 0);
template <typename T> const vec3<T> vec3<T>::UNIT_X(1, 0, 0);
template <typename T> const vec3<T> vec3<T>::UNIT_Y(0, 1, 0);
template <typename T> const vec3<T> vec3<T>::UNIT_Z(0, 0, 1);
typedef vec3<float> vec3f;
typedef vec3<double> vec3d;
}
using namespace ork;
#endif
