/*
 * Plain-C client of the libsr4d C ABI (include/sr4d.h): no Python, no torch.
 * Creates a 4DFlowNet engine (the replacement of predictor.prepare_network, predictor.py:11-29), fills the flat
 * weight buffer with a deterministic pattern through the parameter table, runs one forward pass on synthetic 8^3
 * patches and prints a checksum of the (B,16,16,16,3) prediction; tests/test_gpu_c_client.py runs the same weights and
 * inputs through the Python mirror and compares.
 *
 * Build:  gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_forward.c -o examples/c_abi_forward \
 *             -L /usr/local/cuda/lib64 -lcudart -ldl
 * Run:    examples/c_abi_forward 4dflownet_b200/libsr4d.so
 */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sr4d.h"

#define LOAD(sym) do { *(void**)(&p_##sym) = dlsym(lib, #sym); if (!p_##sym) { fprintf(stderr, "missing symbol %s\n", #sym); return 2; } } while (0)

static float pattern(long long i, float scale) {             /* cheap deterministic pseudo-random in [-scale, scale] */
    unsigned long long x = (unsigned long long)i * 6364136223846793005ULL + 1442695040888963407ULL;
    x ^= x >> 33;
    return scale * ((float)(x % 20001ULL) / 10000.0f - 1.0f);
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "4dflownet_b200/libsr4d.so";
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    int (*p_sr4d_create)(sr4d_t**, int, int, int, int, int, int, int);
    void (*p_sr4d_destroy)(sr4d_t*);
    const char* (*p_sr4d_last_error)(const sr4d_t*);
    int (*p_sr4d_num_tensors)(const sr4d_t*);
    long long (*p_sr4d_flat_size)(const sr4d_t*);
    int (*p_sr4d_param_table)(const sr4d_t*, sr4d_tensor_desc*, int);
    float* (*p_sr4d_params)(sr4d_t*);
    int (*p_sr4d_params_changed)(sr4d_t*, void*);
    int (*p_sr4d_forward)(sr4d_t*, const float*, const float*, const float*, const float*, const float*, const float*,
                          float*, int, void*);
    LOAD(sr4d_create); LOAD(sr4d_destroy); LOAD(sr4d_last_error); LOAD(sr4d_num_tensors); LOAD(sr4d_flat_size);
    LOAD(sr4d_param_table); LOAD(sr4d_params); LOAD(sr4d_params_changed); LOAD(sr4d_forward);

    const int P = 8, R = 2, LOW = 1, HI = 1, B = 2, H = P * R;
    sr4d_t* h = NULL;
    int rc = p_sr4d_create(&h, P, R, LOW, HI, B, 0, 0);
    if (rc != SR4D_OK) { fprintf(stderr, "sr4d_create failed: %d (no sm_100 GPU?)\n", rc); return 1; }

    /* weights: pattern per tensor element, scale = the Keras glorot limit for kernels, 0.02 for biases */
    const int nt = p_sr4d_num_tensors(h);
    sr4d_tensor_desc* tab = (sr4d_tensor_desc*)malloc(sizeof(sr4d_tensor_desc) * nt);
    p_sr4d_param_table(h, tab, nt);
    const long long flat = p_sr4d_flat_size(h);
    float* hw = (float*)calloc((size_t)flat, sizeof(float));
    for (int t = 0; t < nt; ++t) {
        float scale = 0.02f;
        if (tab[t].is_kernel) {
            const int k3 = tab[t].shape[0] * tab[t].shape[1] * tab[t].shape[2];
            scale = sqrtf(6.0f / (float)(k3 * tab[t].shape[3] + k3 * tab[t].shape[4]));
        }
        for (long long i = 0; i < tab[t].count; ++i) hw[tab[t].offset + i] = pattern(1000003LL * t + i, scale);
    }
    cudaMemcpy(p_sr4d_params(h), hw, (size_t)flat * sizeof(float), cudaMemcpyHostToDevice);
    p_sr4d_params_changed(h, NULL);

    /* inputs: u,v,w in [-1,1], magnitudes in [0,0.016] */
    const size_t nin = (size_t)B * P * P * P, nout = (size_t)B * H * H * H * 3;
    float* hin = (float*)malloc(6 * nin * sizeof(float));
    for (int c = 0; c < 6; ++c)
        for (size_t i = 0; i < nin; ++i)
            hin[c * nin + i] = c < 3 ? pattern(7 + 31LL * c + 6LL * (long long)i, 1.0f)
                                     : 0.008f * (1.0f + pattern(11 + 17LL * c + 6LL * (long long)i, 1.0f));
    float *din, *dout;
    cudaMalloc((void**)&din, 6 * nin * sizeof(float));
    cudaMalloc((void**)&dout, nout * sizeof(float));
    cudaMemcpy(din, hin, 6 * nin * sizeof(float), cudaMemcpyHostToDevice);
    rc = p_sr4d_forward(h, din, din + nin, din + 2 * nin, din + 3 * nin, din + 4 * nin, din + 5 * nin, dout, B, NULL);
    if (rc != SR4D_OK) { fprintf(stderr, "sr4d_forward: %d %s\n", rc, p_sr4d_last_error(h)); return 1; }
    float* hout = (float*)malloc(nout * sizeof(float));
    if (cudaMemcpy(hout, dout, nout * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "copy back failed\n"); return 1; }
    double sum = 0, asum = 0;
    for (size_t i = 0; i < nout; ++i) { sum += hout[i]; asum += fabs(hout[i]); }
    printf("tensors %d flat %lld out %zu sum %.9e abssum %.9e first %.9e last %.9e\n", nt, flat, nout, sum, asum,
           (double)hout[0], (double)hout[nout - 1]);

    /* error behaviour: a batch larger than max_batch is rejected with SR4D_EINVAL, nothing is thrown */
    rc = p_sr4d_forward(h, din, din, din, din, din, din, dout, B + 1, NULL);
    printf("oversized batch -> rc %d (%s)\n", rc, p_sr4d_last_error(h));
    cudaFree(din); cudaFree(dout);
    p_sr4d_destroy(h);
    free(hw); free(hin); free(hout); free(tab);
    return rc == SR4D_EINVAL ? 0 : 1;
}
