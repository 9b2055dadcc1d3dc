/*
 * ray.h — drop-in replacement for the header that `futhark <backend> --library ray.fut` generates
 * (reference build rule: futhark/Makefile:26-27; the generated ray.h/ray.c are git-ignored there,
 * futhark/.gitignore:1-2).  The reference driver futhark/main.c (#include "ray.h", main.c:8)
 * compiles unmodified against this file and links against libray_b200.so instead of Futhark's ray.o
 * (futhark/Makefile:20-24).
 *
 * Behind this ABI sit hand-written sm_100a CUDA kernels (raytracers_b200/csrc/).  There is no CPU
 * fallback: futhark_context_new fails (returns a context whose futhark_context_get_error is
 * non-NULL) when no CUDA device is usable.
 *
 * Each declaration cites the reference call site it serves.
 */
#ifndef RAY_B200_FUTHARK_COMPAT_RAY_H
#define RAY_B200_FUTHARK_COMPAT_RAY_H

#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUTHARK_BACKEND_cuda
#define RAY_B200 1

/* ---- Initialisation (main.c:59-64, 140-141) ------------------------------------------------ */
struct futhark_context_config;
struct futhark_context_config *futhark_context_config_new(void);                 /* main.c:59 */
void futhark_context_config_free(struct futhark_context_config *cfg);            /* main.c:141 */
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_cache_file(struct futhark_context_config *cfg, const char *f);
/* CUDA-backend style device selection: "#k" or "k" picks device k. */
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s);
/* Tuning parameters: see futhark_get_tuning_param_name().  Returns 0 on success, 1 if unknown. */
int futhark_context_config_set_tuning_param(struct futhark_context_config *cfg, const char *param_name,
                                            size_t new_value);
int futhark_get_tuning_param_count(void);
const char *futhark_get_tuning_param_name(int i);
const char *futhark_get_tuning_param_class(int i);

struct futhark_context;
struct futhark_context *futhark_context_new(struct futhark_context_config *cfg); /* main.c:61 */
void futhark_context_free(struct futhark_context *ctx);                          /* main.c:140 */

/* ---- Arrays: [h][w]i32 image (ray.fut:164, 246-247) ----------------------------------------- */
struct futhark_i32_2d;
struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t dim0,
                                          int64_t dim1);
/* Wraps (does not copy, does not own) a device pointer. */
struct futhark_i32_2d *futhark_new_raw_i32_2d(struct futhark_context *ctx, void *device_ptr, int64_t dim0,
                                              int64_t dim1);
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);          /* main.c:110,137 */
/* Copies the image to caller memory; the copy has completed when this returns (main.c:130 does
 * not sync afterwards). */
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr, int32_t *data); /* main.c:130 */
/* Device pointer of the row-major image (lets a caller keep the frame on the GPU). */
void *futhark_values_raw_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);
const int64_t *futhark_shape_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);

/* ---- Opaque values (ray.fut:171-174 `scene`, ray.fut:239 `prepared_scene`) ------------------- */
struct futhark_opaque_scene;
int futhark_free_opaque_scene(struct futhark_context *ctx, struct futhark_opaque_scene *obj);   /* main.c:139 */
int futhark_store_opaque_scene(struct futhark_context *ctx, const struct futhark_opaque_scene *obj, void **p,
                               size_t *n);
struct futhark_opaque_scene *futhark_restore_opaque_scene(struct futhark_context *ctx, const void *p);

struct futhark_opaque_prepared_scene;
int futhark_free_opaque_prepared_scene(struct futhark_context *ctx,
                                       struct futhark_opaque_prepared_scene *obj);              /* main.c:91,138 */
int futhark_store_opaque_prepared_scene(struct futhark_context *ctx,
                                        const struct futhark_opaque_prepared_scene *obj, void **p, size_t *n);
struct futhark_opaque_prepared_scene *futhark_restore_opaque_prepared_scene(struct futhark_context *ctx,
                                                                            const void *p);

/* ---- Entry points (ray.fut:176, 223, 241, 246) ---------------------------------------------- */
int futhark_entry_rgbbox(struct futhark_context *ctx, struct futhark_opaque_scene **out0);      /* main.c:73 */
int futhark_entry_irreg(struct futhark_context *ctx, struct futhark_opaque_scene **out0);       /* main.c:76 */
/* prepare_scene h w scene (ray.fut:241-244): note the order (height, width) — main.c:94-96. */
int futhark_entry_prepare_scene(struct futhark_context *ctx, struct futhark_opaque_prepared_scene **out0,
                                const int64_t in0, const int64_t in1, const struct futhark_opaque_scene *in2);
/* render h w prepared_scene (ray.fut:246-247) — main.c:113-115.  THE HOT PATH.  Asynchronous:
 * completion is guaranteed after futhark_context_sync (main.c:117). */
int futhark_entry_render(struct futhark_context *ctx, struct futhark_i32_2d **out0, const int64_t in0,
                         const int64_t in1, const struct futhark_opaque_prepared_scene *in2);

/* ---- Miscellaneous --------------------------------------------------------------------------- */
int futhark_context_sync(struct futhark_context *ctx);                                          /* main.c:98,117 */
/* Returns NULL if no error is pending, else a malloc'ed string the caller frees (main.c:64). */
char *futhark_context_get_error(struct futhark_context *ctx);
/* malloc'ed human-readable report (kernel timings when profiling is on); caller frees. */
char *futhark_context_report(struct futhark_context *ctx);
void futhark_context_set_logging_file(struct futhark_context *ctx, FILE *f);
void futhark_context_pause_profiling(struct futhark_context *ctx);
void futhark_context_unpause_profiling(struct futhark_context *ctx);
int futhark_context_clear_caches(struct futhark_context *ctx);

#ifdef __cplusplus
}
#endif

#endif /* RAY_B200_FUTHARK_COMPAT_RAY_H */
