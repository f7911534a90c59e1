/* Shim standing in for Ork's ork/core/Object.h (Ork is not vendored in the reference tree) so that the reference's
 * preprocess/terrain sources compile UNCHANGED from where they lie under /root/reference.  Test infrastructure.
 * What those sources take from it: <cassert>/<cstdio>, namespace std, and Ork's portable fopen(FILE **, name, mode). */
#ifndef ORC_SHIM_ORK_OBJECT_H
#define ORC_SHIM_ORK_OBJECT_H
#include <cassert>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <exception>
#include <string>
#include <unistd.h>
#include <sys/stat.h>
#include <sys/types.h>
#ifndef PROLAND_API
#define PROLAND_API
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
using namespace std;
/* Util.cpp's flog() keeps a one-line journal "log.txt" in the working directory; the shim keeps it out of the tree:
 * reading finds nothing, writing goes to an anonymous temporary file */
inline void fopen(FILE **f, const char *name, const char *mode)
{
    if (strcmp(name, "log.txt") == 0) *f = mode[0] == 'r' ? NULL : tmpfile();
    else *f = ::fopen(name, mode);
}
/* Ork's 64-bit seek */
inline int fseek64(FILE *f, long long off, int whence) { return fseeko(f, (off_t) off, whence); }
/* the reference is built for MinGW, whose mkdir takes one argument */
inline int mkdir(const char *path) { return ::mkdir(path, 0777); }
#endif
