/* Minimal C consumer of the C ABI (include/cleanba_b200.h): no Python, no torch -- plain pointers and sizes.
 *
 *   gcc -std=c99 -I include examples/abi_smoke.c -o /tmp/abi_smoke -ldl && /tmp/abi_smoke cleanba_b200/libcleanba_b200.so
 *
 * On a B200 it creates an actor context, uploads zero parameters, runs one cb_actor_step on 8 zero frames and prints the actions
 * (all logits are 0 with zero parameters, so the actions are the Gumbel-max draws of key (0, 1)).  On a machine without a GPU
 * cb_create fails and the program prints cb_last_error() and exits with status 2: the error path of the ABI needs no device.
 * The CUDA runtime calls (cudaMalloc / cudaMemcpy) are resolved from the same process image with dlsym, so the example has no
 * build-time dependency on the CUDA toolkit. */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cleanba_b200.h"

typedef int (*create_fn)(const cb_config*, cb_ctx**);
typedef void (*destroy_fn)(cb_ctx*);
typedef const char* (*err_fn)(void);
typedef long long (*nparam_fn)(int);
typedef int (*set_params_fn)(cb_ctx*, const float*, cb_stream);
typedef int (*actor_fn)(cb_ctx*, const uint8_t*, int, uint32_t*, int32_t*, float*, float*, float*, cb_stream);
typedef int (*cuda_malloc_fn)(void**, size_t);
typedef int (*cuda_memcpy_fn)(void*, const void*, size_t, int);
typedef int (*cuda_memset_fn)(void*, int, size_t);
typedef int (*cuda_sync_fn)(void);

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "cleanba_b200/libcleanba_b200.so";
    void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "dlopen failed: %s\n", dlerror()); return 1; }
    create_fn cb_create_p = (create_fn)dlsym(h, "cb_create");
    destroy_fn cb_destroy_p = (destroy_fn)dlsym(h, "cb_destroy");
    err_fn cb_last_error_p = (err_fn)dlsym(h, "cb_last_error");
    nparam_fn cb_num_params_p = (nparam_fn)dlsym(h, "cb_num_params");
    set_params_fn cb_set_params_p = (set_params_fn)dlsym(h, "cb_set_params");
    actor_fn cb_actor_step_p = (actor_fn)dlsym(h, "cb_actor_step");
    if (!cb_create_p || !cb_destroy_p || !cb_last_error_p || !cb_num_params_p || !cb_set_params_p || !cb_actor_step_p) {
        fprintf(stderr, "missing symbol: %s\n", dlerror());
        return 1;
    }
    printf("parameters: %lld floats\n", cb_num_params_p(18));

    cb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0; cfg.algo = CB_ALGO_PPO; cfg.max_batch = 8; cfg.train = 0; cfg.num_actions = 18; cfg.conv_backend = CB_CONV_TCGEN05;
    cb_ctx* ctx = NULL;
    if (cb_create_p(&cfg, &ctx) != 0) {
        printf("cb_create failed (expected without a B200): %s\n", cb_last_error_p());
        return 2;
    }
    cuda_malloc_fn cu_malloc = (cuda_malloc_fn)dlsym(RTLD_DEFAULT, "cudaMalloc");
    cuda_memcpy_fn cu_memcpy = (cuda_memcpy_fn)dlsym(RTLD_DEFAULT, "cudaMemcpy");
    cuda_memset_fn cu_memset = (cuda_memset_fn)dlsym(RTLD_DEFAULT, "cudaMemset");
    cuda_sync_fn cu_sync = (cuda_sync_fn)dlsym(RTLD_DEFAULT, "cudaDeviceSynchronize");
    if (!cu_malloc || !cu_memcpy || !cu_memset || !cu_sync) { fprintf(stderr, "CUDA runtime symbols not found\n"); return 1; }
    const int n = 8;
    const long long np = cb_num_params_p(18);
    float* params = (float*)calloc((size_t)np, sizeof(float));
    void *obs = NULL, *key = NULL, *action = NULL, *logprob = NULL, *value = NULL;
    cu_malloc(&obs, (size_t)n * 4 * 84 * 84); cu_memset(obs, 0, (size_t)n * 4 * 84 * 84);
    cu_malloc(&key, 8); cu_malloc(&action, n * 4); cu_malloc(&logprob, n * 4); cu_malloc(&value, n * 4);
    const uint32_t k[2] = {0u, 1u};
    cu_memcpy(key, k, 8, 1 /* cudaMemcpyHostToDevice */);
    int rc = cb_set_params_p(ctx, params, NULL);          /* host pointer: cudaMemcpyDefault inside the library */
    if (!rc) rc = cb_actor_step_p(ctx, (const uint8_t*)obs, n, (uint32_t*)key, (int32_t*)action, (float*)logprob, (float*)value, NULL, NULL);
    if (rc) { printf("call failed: %s\n", cb_last_error_p()); cb_destroy_p(ctx); return 3; }
    cu_sync();
    int32_t a[8];
    cu_memcpy(a, action, sizeof(a), 2 /* cudaMemcpyDeviceToHost */);
    printf("actions:");
    for (int i = 0; i < n; ++i) printf(" %d", a[i]);
    printf("\n");
    cb_destroy_p(ctx);
    free(params);
    return 0;
}
