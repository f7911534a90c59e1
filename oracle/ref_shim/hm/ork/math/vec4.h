/* Shim standing in for Ork's ork/math/vec4.h: the preprocess sources only use vec4f as a bag of four floats.
 * Test infrastructure. */
#ifndef ORC_SHIM_ORK_VEC4_H
#define ORC_SHIM_ORK_VEC4_H
#include "ork/core/Object.h"
namespace ork {
template <typename T> struct vec4 {
    T x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(T x, T y, T z, T w) : x(x), y(y), z(z), w(w) {}
};
typedef vec4<float> vec4f;
}
using namespace ork;
#endif
