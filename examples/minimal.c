/* examples/minimal.c — the C ABI from plain C: one plane from an in-memory particle list.
 * cc -std=c99 -Iinclude examples/minimal.c -Lslicer_b200/_build -lslicer_b200 -Wl,-rpath,$PWD/slicer_b200/_build -o minimal
 * (tests/test_capi_abi.py compiles and links this file; running it needs a B200.) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "slicer_b200.h"

int main(void)
{
  const size_t n = 1u << 20;
  const double box = 128000.0; /* kpc/h */
  const int npix = 256;
  slicer_config cfg;
  slicer_handle *h = NULL;
  slicer_plane_desc plane;
  const double massarr[SLICER_NTYPES] = {0, 1.0375, 0, 0, 0, 0};
  float *pos = NULL, *map = NULL;
  long long counts[SLICER_NTYPES], ingrid[SLICER_NTYPES];
  size_t i;

  memset(&cfg, 0, sizeof(cfg));
  cfg.device = 0;
  cfg.mas = SLICER_MAS_TSC;
  cfg.max_m = 1e3; /* MAX_M, densitymaps.h:21 */
  cfg.max_planes = 1;
  cfg.npix_max = npix;
  cfg.particle_capacity = n;
  if (slicer_create(&cfg, &h))
  {
    fprintf(stderr, "slicer_create: %s\n", slicer_last_error());
    return 1;
  }
  if (slicer_alloc_pinned(n * 3 * sizeof(float), (void **)&pos))
    return 1;
  srand(1);
  for (i = 0; i < 3 * n; i++)
    pos[i] = (float)(box * (rand() / (RAND_MAX + 1.0)));

  memset(&plane, 0, sizeof(plane));
  plane.sgn[0] = plane.sgn[1] = plane.sgn[2] = 1;
  plane.face = 1;
  plane.rcase = 1.0f;           /* the box is piled once along the line of sight */
  plane.ld = 160.0;             /* Mpc/h */
  plane.ld2 = 192.0;
  plane.fovradiants = 0.2;
  plane.npix = npix;

  map = (float *)malloc((size_t)npix * npix * sizeof(float));
  if (slicer_begin_snapshot(h, box, massarr, 0) || slicer_stage_particles(h, 1, pos, SLICER_LAYOUT_AOS, NULL, n) ||
      slicer_deposit(h, &plane, 1) || slicer_fetch(h, 0, -1, map, counts, ingrid))
  {
    fprintf(stderr, "%s\n", slicer_last_error());
    return 1;
  }
  printf("accepted %lld particles of type 1, %lld with their nearest grid point inside the map\n", counts[1], ingrid[1]);
  free(map);
  slicer_free_pinned(pos);
  slicer_destroy(h);
  return 0;
}
