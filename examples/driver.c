/*
 * driver.c — a small C consumer of the Futhark-compatible ABI (include/ray.h), written for this repo.
 * It follows the call protocol of the reference driver futhark/main.c:59-141 (context, scene entry,
 * timed prepare_scene loop, timed render loop with futhark_context_sync inside, values, frees) so the
 * drop-in boundary can be exercised on machines that do not have the reference tree.  Differences:
 * prints a checksum of the frame instead of writing a P3 file unless -f is given, and accepts -p SPP.
 *
 *   driver [-s rgbbox|irreg] [-n HEIGHT] [-m WIDTH] [-r RUNS] [-p SPP] [-f out.ppm]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "ray_b200.h"

static double now_s(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      char *e_ = futhark_context_get_error(ctx);                           \
      fprintf(stderr, "%s failed: %s\n", #call, e_ ? e_ : "(no message)"); \
      free(e_);                                                            \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  int height = 200, width = 200, runs = 10, spp = 1;
  const char *scene_name = "rgbbox", *out = NULL;
  for (int a = 1; a + 1 < argc; a += 2) {
    if (!strcmp(argv[a], "-n")) height = atoi(argv[a + 1]);
    else if (!strcmp(argv[a], "-m")) width = atoi(argv[a + 1]);
    else if (!strcmp(argv[a], "-r")) runs = atoi(argv[a + 1]);
    else if (!strcmp(argv[a], "-p")) spp = atoi(argv[a + 1]);
    else if (!strcmp(argv[a], "-s")) scene_name = argv[a + 1];
    else if (!strcmp(argv[a], "-f")) out = argv[a + 1];
    else { fprintf(stderr, "unknown option %s\n", argv[a]); return 2; }
  }
  struct futhark_context_config *cfg = futhark_context_config_new();
  struct futhark_context *ctx = futhark_context_new(cfg);
  char *err = ctx ? futhark_context_get_error(ctx) : NULL;
  if (!ctx || err) { fprintf(stderr, "context: %s\n", err ? err : "NULL"); free(err); return 1; }
  CHECK(ray_b200_context_set_spp(ctx, spp));

  struct futhark_opaque_scene *scene = NULL;
  if (!strcmp(scene_name, "rgbbox")) CHECK(futhark_entry_rgbbox(ctx, &scene));
  else if (!strcmp(scene_name, "irreg")) CHECK(futhark_entry_irreg(ctx, &scene));
  else { fprintf(stderr, "unknown scene %s\n", scene_name); return 2; }

  struct futhark_opaque_prepared_scene *prep = NULL;
  double t0 = now_s();
  for (int i = 0; i < runs; i++) {
    if (prep) CHECK(futhark_free_opaque_prepared_scene(ctx, prep));
    CHECK(futhark_entry_prepare_scene(ctx, &prep, height, width, scene));
    CHECK(futhark_context_sync(ctx));
  }
  printf("prepare_scene: %.6f s/run\n", (now_s() - t0) / runs);

  struct futhark_i32_2d *img = NULL;
  t0 = now_s();
  for (int i = 0; i < runs; i++) {
    if (img) CHECK(futhark_free_i32_2d(ctx, img));
    CHECK(futhark_entry_render(ctx, &img, height, width, prep));
    CHECK(futhark_context_sync(ctx));
  }
  printf("render: %.6f s/run\n", (now_s() - t0) / runs);

  int32_t *host = malloc(sizeof(int32_t) * (size_t)height * (size_t)width);
  CHECK(futhark_values_i32_2d(ctx, img, host));
  unsigned long long sum = 1469598103934665603ULL; /* FNV-1a over the packed pixels */
  for (long k = 0; k < (long)height * width; k++) { sum ^= (unsigned)host[k]; sum *= 1099511628211ULL; }
  printf("frame %dx%d fnv1a=%016llx first=%06x\n", width, height, sum, (unsigned)host[0]);
  if (out) {
    FILE *f = fopen(out, "w");
    if (!f) { perror(out); return 1; }
    fprintf(f, "P3\n%d %d\n255\n", width, height);
    for (long k = 0; k < (long)height * width; k++)
      fprintf(f, "%d %d %d\n", (host[k] >> 16) & 0xFF, (host[k] >> 8) & 0xFF, host[k] & 0xFF);
    fclose(f);
  }
  free(host);
  CHECK(futhark_free_i32_2d(ctx, img));
  CHECK(futhark_free_opaque_prepared_scene(ctx, prep));
  CHECK(futhark_free_opaque_scene(ctx, scene));
  futhark_context_free(ctx);
  futhark_context_config_free(cfg);
  return 0;
}
