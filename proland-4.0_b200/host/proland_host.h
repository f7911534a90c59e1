/*
 * proland_host.h -- flat C view of the C++ host layer (namespace proland:
 * TileStorage / TileCache / TileProducer / Elevation-, Normal-, ResidualProducer /
 * ResourceManager / BatchScheduler), used by the Python tests, bench.py and tools
 * through ctypes.  Handles are opaque pointers to the C++ objects.  Functions
 * returning int give 0 on success, -1 on error (text in plh_last_error());
 * functions returning a handle give NULL on error.
 */
#ifndef PROLAND_HOST_H
#define PROLAND_HOST_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *plh_last_error(void);

/* resources of an XML archive (ResourceManager) */
void *plh_open(const char *xml, const char *data_dir, int device);
void plh_close(void *mgr);
void plh_shutdown(void);
void *plh_producer(void *mgr, const char *name);
void *plh_cache(void *mgr, const char *name);
void *plh_scheduler(void *mgr, const char *name);
void *plh_producer_cache(void *prod);
void *plh_cache_scheduler(void *cache);

/* TileProducer */
void plh_set_root_quad_size(void *prod, float size);
int plh_producer_info(void *prod, int out[6]);   /* id, border, gpu, tile size, #referenced producers, root quad size */
const char *plh_producer_type(void *prod);
const char *plh_producer_task_type(void *prod);
int plh_has_tile(void *prod, int level, int tx, int ty);
int plh_has_children(void *prod, int level, int tx, int ty);
void *plh_get_tile(void *prod, int level, int tx, int ty, unsigned int deadline);
void *plh_find_tile(void *prod, int level, int tx, int ty, int include_cache, int done);
int plh_put_tile(void *prod, void *tile);
int plh_prefetch_tile(void *prod, int level, int tx, int ty);
void plh_invalidate_tiles(void *prod);
void plh_invalidate_tile(void *prod, int level, int tx, int ty);
int plh_producer_counts(void *prod, unsigned long long out[2]);   /* tiles made, batches launched */

/* TileCache::Tile */
int plh_tile_done(void *tile);
int plh_tile_slot(void *tile);
int plh_tile_download(void *tile, void *buf, size_t bytes);
int plh_tile_minmax(void *prod, void *tile, float out[2]);

/* Scheduler::run over the tasks of n tiles */
int plh_run(void *scheduler, void **tiles, int n);
int plh_cache_stats(void *cache, int out[6]);   /* used, unused, capacity, free slots, queries, misses */
int plh_scheduler_stats(void *scheduler, unsigned long long out[4]);   /* frame, waves, tasks, queued prefetches */

/* ResidualProducer file arithmetic */
int plh_residual_info(void *prod, int out[3]);   /* minLevel, maxLevel, deltaLevel */
int plh_residual_tile_id(void *prod, int level, int tx, int ty);
int plh_residual_tile_size(void *prod, int level);

unsigned long long plh_device_launches(int device);
int plh_device_sync(int device);
void plh_debug_log(int on, int echo);
unsigned long plh_debug_log_lines(void);
void plh_quiet_errors(int quiet);
int plh_upsample_variant(const char *prog, int out[2]);

/* TerrainNode / TerrainQuad (view-dependent quadtree) and TileSampler (which tiles the quads need) */
void *plh_terrain_create(float size, float zmin, float zmax, float split_factor, int max_level);
void plh_terrain_destroy(void *node);
float plh_split_distance(float split_factor, float viewport_width, float fov_radians);
int plh_terrain_update(void *node, double x, double y, double z, float split_dist, float dist_factor);
int plh_terrain_quads(void *node, int *out, int max_quads);   /* pre-order (level, tx, ty, leaf) */
void *plh_sampler_create(const char *name, void *prod, int async, int store_parent);
void *plh_sampler_z_create(const char *name, void *prod, int async, int store_parent);
int plh_sampler_z_counts(void *sampler, unsigned long long out[3]);   /* read-backs issued, applied, tiles waiting */
void plh_ground_height(float out[2], int reset);   /* TerrainNode::groundHeightAtCamera, nextGroundHeightAtCamera */
int plh_terrain_quads_z(void *node, float *out, int max_quads);   /* pre-order (level, tx, ty, zmin, zmax) */
void plh_sampler_destroy(void *sampler);
int plh_sampler_tile_count(void *sampler);
int plh_frame_update(void *scheduler, void *node, void **samplers, int n);
/* proland::preprocessDem / preprocessSphericalDem (Preprocess.cpp:512-585) over a float array: writes dst_folder/DEM.dat
 * or DEM1..6.dat; 0 or -1 (plh_last_error) */
int plh_preprocess_dem(const float *src, int src_w, int src_h, int min_tile_size, int tile_size, int max_level, const char *dst_folder,
                       float residual_scale, int spherical);

/* CPU-only scene with a recording producer (cache / task / scheduler logic without a device) */
void *plh_test_scene(int capacity, int tile_size, int max_level, int prefetch_rate, int prefetch_queue);
void plh_test_scene_close(void *scene);
void *plh_test_producer(void *scene);
void *plh_test_cache(void *scene);
void *plh_test_scheduler(void *scene);
int plh_test_calls(void *scene, int *out, int max_calls);
int plh_test_begin_end(void *scene, int out[2]);

#ifdef __cplusplus
}
#endif
#endif
