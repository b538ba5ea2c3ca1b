/* Minimal C host of libh2o_b200.so: proves the ABI is plain C (no C++/torch types) and shows the
 * call sequence a non-Python embedder would use.  Build:
 *   gcc -std=c99 -Iinclude examples/c_host.c -Lsilver2_isaacsim_b200/lib -lh2o_b200 -lcudart -o c_host
 * Device buffers are allocated with the CUDA runtime; everything else is the header's API. */
#include <stdio.h>
#include <stdlib.h>

#include "h2o.h"
#include "h2o_dlpack.h"

extern int cudaMalloc(void** p, size_t n);
extern int cudaMemset(void* p, int v, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);

int main(void)
{
    const int64_t n = 1024;
    h2o_handle h = NULL;
    if (h2o_create(&h, n, H2O_F32, 0) != H2O_OK) {
        fprintf(stderr, "h2o_create: %s\n", h2o_last_error());
        return 1; /* e.g. no GPU: the engine has no CPU path */
    }
    /* README-default 1 m cube, reference wrapper ctor order (numba_hydrodynamics_wrapper.py:9-10) */
    const double ctor[12] = {1, 1, 1, 1.2, 0.8, 300, 150, 1025, 9.81, 0.05, 0.02, 1.0};
    if (h2o_set_params_uniform(h, ctor, 512.5)) return 2;
    float *pos, *quat, *lin, *ang, *F, *T;
    cudaMalloc((void**)&pos, n * 3 * 4); cudaMalloc((void**)&quat, n * 4 * 4);
    cudaMalloc((void**)&lin, n * 3 * 4); cudaMalloc((void**)&ang, n * 3 * 4);
    cudaMalloc((void**)&F, n * 3 * 4);   cudaMalloc((void**)&T, n * 3 * 4);
    float* hq = (float*)calloc(n * 4, 4);
    float* hp = (float*)calloc(n * 3, 4);
    for (int64_t i = 0; i < n; ++i) { hq[4 * i + 3] = 1.0f; hp[3 * i + 2] = -0.2f; }
    cudaMemcpy(quat, hq, n * 16, 1); cudaMemcpy(pos, hp, n * 12, 1);
    cudaMemset(lin, 0, n * 12); cudaMemset(ang, 0, n * 12);
    int rc = h2o_step(h, pos, quat, lin, ang, 1.0 / 60.0, F, T, NULL, NULL);
    if (rc) { fprintf(stderr, "h2o_step: %s\n", h2o_last_error()); return 3; }
    cudaDeviceSynchronize();
    float f0[3];
    cudaMemcpy(f0, F, 12, 2);
    printf("%s: F[0] = (%g, %g, %g) N  (buoyancy of a 70%% submerged 1 m^3 cube: 7038.67 N)\n", h2o_version(), f0[0], f0[1], f0[2]);
    h2o_destroy(h);
    return 0;
}
