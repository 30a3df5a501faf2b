/*
 * Plain-C client of the C ABI (include/mr_rl_b200.h): no Python, no torch — device buffers from the CUDA
 * runtime, one env, noise-free, a few MR_Env.step calls.  Prints the positions so the test can compare
 * them with the golden vectors of the live reference.  With MR_DEMO_HOST=1 in the environment every step goes through
 * mr_env_step_host in direct mode instead: actions and results live in page-locked HOST buffers (cudaHostAlloc) that the
 * kernel reads and writes itself — the call a numpy-holding caller of the reference would make.  Build:
 *   gcc c_abi_demo.c -I../include -I/usr/local/cuda/include -L../mr_rl_b200/_lib -lmr_rl_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -o c_abi_demo
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mr_rl_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "cuda: %s\n", cudaGetErrorString(e_)); return 2; } } while (0)
#define MR(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "mr: %d %s\n", r_, mr_last_error()); return 3; } } while (0)

int main(int argc, char** argv) {
    /* argv: x0 y0 a0 n_steps f_0 alpha_0 f_1 alpha_1 ... */
    if (argc < 5) return 1;
    const double x0 = atof(argv[1]), y0 = atof(argv[2]), a0 = atof(argv[3]);
    const int steps = atoi(argv[4]);
    const int64_t n = 1, pad = 16;
    double *state, *obs, *rew, *act, *tt_dev, *init;
    int32_t *counter, *cursor;
    uint8_t *status, *done;
    CK(cudaMalloc((void**)&state, 5 * pad * sizeof(double)));
    CK(cudaMalloc((void**)&obs, 5 * pad * sizeof(double)));
    CK(cudaMalloc((void**)&rew, pad * sizeof(double)));
    CK(cudaMalloc((void**)&act, 2 * pad * sizeof(double)));
    CK(cudaMalloc((void**)&init, 2 * pad * sizeof(double)));
    CK(cudaMalloc((void**)&counter, pad * sizeof(int32_t)));
    CK(cudaMalloc((void**)&cursor, pad * sizeof(int32_t)));
    CK(cudaMalloc((void**)&status, pad));
    CK(cudaMalloc((void**)&done, pad));
    CK(cudaMemset(status, 0, pad));
    double tt_host[256];
    mr_fill_time_table_host(tt_host, 256, 0.030);
    CK(cudaMalloc((void**)&tt_dev, sizeof(tt_host)));
    CK(cudaMemcpy(tt_dev, tt_host, sizeof(tt_host), cudaMemcpyHostToDevice));

    mr_sim_params p;
    mr_default_params(&p);
    p.a0 = a0; p.noise_var = 0.0;                       /* sigma = 0: no noise source needed */
    mr_env_state st = {state, state + pad, state + 2 * pad, state + 3 * pad, state + 4 * pad, counter, cursor, status};
    mr_time_table tt = {tt_dev, 256, 0};
    mr_step_out out = {obs, rew, done, NULL, pad};
    const double xy[2] = {x0, y0};
    CK(cudaMemcpy(init, xy, sizeof(xy), cudaMemcpyHostToDevice));
    MR(mr_env_reset(&st, n, MR_F64, &p, NULL, init, NULL, 1, &out, NULL));
    const char* host_mode = getenv("MR_DEMO_HOST");
    if (host_mode && host_mode[0] == '1') {
        double *h_act, *h_obs, *h_rew;
        uint8_t* h_done;
        CK(cudaHostAlloc((void**)&h_act, 2 * pad * sizeof(double), cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&h_obs, 5 * pad * sizeof(double), cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&h_rew, pad * sizeof(double), cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&h_done, pad, cudaHostAllocDefault));
        memset(h_obs, 0, 5 * pad * sizeof(double));          /* the goal rows stay zero: they are never re-sent */
        mr_host_step_io io = {h_act, NULL, h_obs, h_rew, h_done, pad, 0, 0};
        for (int k = 0; k < steps; ++k) {
            h_act[0] = atof(argv[5 + 2 * k]); h_act[1] = atof(argv[6 + 2 * k]);
            MR(mr_env_step_host(NULL, &st, n, MR_F64, &p, NULL, &tt, &io, NULL, 0, NULL));   /* blocks until filled */
            printf("%.17g %.17g %.17g %.17g %d\n", h_obs[0], h_obs[pad], h_obs[4 * pad], h_rew[0], (int)h_done[0]);
        }
        if (h_obs[2 * pad] != 0.0 || h_obs[3 * pad] != 0.0) return 5;
        if (mr_env_step_host(NULL, &st, n, MR_F64, &p, NULL, &tt, NULL, NULL, 0, NULL) != MR_ERR_ARG) return 4;
        return 0;
    }
    for (int k = 0; k < steps; ++k) {
        const double a[2] = {atof(argv[5 + 2 * k]), atof(argv[6 + 2 * k])};
        CK(cudaMemcpy(act, a, sizeof(a), cudaMemcpyHostToDevice));
        MR(mr_env_step(&st, n, MR_F64, &p, NULL, &tt, act, &out, NULL));
        double o[5]; double r; uint8_t d;
        CK(cudaMemcpy(&o[0], obs, sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&o[1], obs + pad, sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&o[4], obs + 4 * pad, sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&r, rew, sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&d, done, 1, cudaMemcpyDeviceToHost));
        printf("%.17g %.17g %.17g %.17g %d\n", o[0], o[1], o[4], r, (int)d);
    }
    /* error path: null actions is reported, not thrown */
    if (mr_env_step(&st, n, MR_F64, &p, NULL, &tt, NULL, &out, NULL) != MR_ERR_ARG) return 4;
    return 0;
}
