// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/cu/lvtime scripts/cu/lvtime.cu emloco_b200/libemloco_b200.so -Xlinker -rpath -Xlinker '$ORIGIN/../../emloco_b200'
// standalone timing of emloco_locoval_forward through the C ABI: cudaEvent pairs and %globaltimer stamps
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/emloco.h"
__global__ void stamp(unsigned long long* t) { unsigned long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); *t = v; }
int main(int argc, char** argv) {
    long long B = argc > 1 ? atoll(argv[1]) : (1 << 20);
    std::vector<float> traj(B * 26), pose(B * 72), vel(B * 2), w(6174);
    srand(1);
    auto fill = [](std::vector<float>& v, float s) { for (auto& x : v) x = s * ((float)rand() / RAND_MAX - 0.5f); };
    fill(traj, 2.f); fill(pose, 1.f); fill(vel, 2.f); fill(w, 0.3f);
    float *dt, *dp, *dv, *dw, *dval; unsigned long long* ts;
    cudaMalloc(&dt, traj.size() * 4); cudaMalloc(&dp, pose.size() * 4); cudaMalloc(&dv, vel.size() * 4); cudaMalloc(&dw, w.size() * 4);
    cudaMalloc(&dval, B * 4); cudaMalloc(&ts, 64 * 8);
    cudaMemcpy(dt, traj.data(), traj.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dp, pose.data(), pose.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, vel.data(), vel.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    int flags = 1 | 2 | 4 | 8 | 16;
    for (int i = 0; i < 3; ++i) if (emloco_locoval_forward(dt, 2, 13, dp, dv, dw, dval, B, flags, 0)) { printf("err %s\n", emloco_last_error()); return 1; }
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); emloco_locoval_forward(dt, 2, 13, dp, dv, dw, dval, B, flags, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); printf("single launch, events: %.1f us\n", ms * 1e3);
    }
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) emloco_locoval_forward(dt, 2, 13, dp, dv, dw, dval, B, flags, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("20 back to back, events: %.1f us each\n", ms * 1e3 / 20);
    for (int i = 0; i < 8; ++i) { stamp<<<1, 1>>>(ts + 2 * i); emloco_locoval_forward(dt, 2, 13, dp, dv, dw, dval, B, flags, 0); stamp<<<1, 1>>>(ts + 2 * i + 1); }
    cudaDeviceSynchronize();
    unsigned long long h[16]; cudaMemcpy(h, ts, sizeof h, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i) printf("globaltimer stamps around launch %d: %.1f us\n", i, (h[2 * i + 1] - h[2 * i]) * 1e-3);
    return 0;
}
