/* CPU-only PROFILING stub of the C ABI (never shipped, never loaded by the product or the tests): every device entry point
 * succeeds and does nothing, so that gprof sees the pure host cost of the plugin path. */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
typedef struct pl_ctx { int d; unsigned long long launches; } pl_ctx;
typedef struct pl_pool { int kind, w, cap; } pl_pool;
const char *pl_last_error(void) { return "stub"; }
int pl_ctx_create(int device, pl_ctx **out) { *out = calloc(1, sizeof(pl_ctx)); return 0; }
void pl_ctx_destroy(pl_ctx *c) { free(c); }
unsigned long long pl_ctx_launch_count(pl_ctx *c) { return c->launches; }
int pl_sync(pl_ctx *c) { return 0; }
int pl_noise_init(pl_ctx *c, int w, void *p) { return 0; }
int pl_ortho_noise_init(pl_ctx *c, int w, void *p) { return 0; }
int pl_pool_create(pl_ctx *c, int kind, int w, int cap, pl_pool **out) { pl_pool *p = calloc(1, sizeof(pl_pool)); p->kind = kind; p->w = w; p->cap = cap; *out = p; return 0; }
void pl_pool_destroy(pl_pool *p) { free(p); }
int pl_pool_capacity(const pl_pool *p) { return p->cap; }
size_t pl_pool_tile_bytes(const pl_pool *p) { return (size_t) p->w * p->w * 4; }
int pl_pool_download(pl_pool *p, int s, void *h, size_t b) { return 0; }
int pl_pool_upload(pl_pool *p, int s, const void *h, size_t b) { return 0; }
int pl_elevation_batch() { return 0; }
int pl_normal_batch() { return 0; }
int pl_ortho_batch() { return 0; }
int pl_ortho_decode_batch() { return 0; }
int pl_residual_decode_batch() { return 0; }
int pl_residual_encode_batch() { return 0; }
int pl_residual_upsample() { return 0; }
int pl_residual_write_file() { return 0; }
int pl_elev_stats_download() { return 0; }
int pl_elev_stats_readback_end() { return 0; }
int pl_elev_zreadback_begin() { return 0; }
int pl_height_cube_from_latlon() { return 0; }
int pl_height_cube_from_plane() { return 0; }
int pl_height_tiles() { return 0; }
void pl_height_cube_destroy() {}
void pl_ortho_make_req() {}
