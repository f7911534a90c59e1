/* Shim standing in for Ork's pmath.h (Ork is not vendored in the reference
 * tree) so that the reference's core/sources/proland/math/noise.cpp compiles
 * UNCHANGED from where it lies under /root/reference.  Test infrastructure. */
#ifndef ORC_SHIM_PMATH_H
#define ORC_SHIM_PMATH_H
#include <cmath>
#include <cstdlib>
using std::sqrt;
using std::log;
using std::floor;
#ifndef PROLAND_API
#define PROLAND_API
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#endif
