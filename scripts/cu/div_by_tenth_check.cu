#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float div_by_tenth(float x) {
    const float q0 = __fmul_rn(x, 10.0f);
    return __fmaf_rn(__fmaf_rn(-q0, 0.1f, x), 10.0f, q0);
}
__global__ void k(unsigned long long* bad, unsigned long long* badint) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = i; b < 0x4B800000ull; b += stride) {   // all positive floats up to 2^24
        float x = __uint_as_float((unsigned)b);
        float a = __fdiv_rn(x, 0.1f), c = div_by_tenth(x);
        if (a != c) { atomicAdd(bad, 1ull); if ((long long)a != (long long)c) atomicAdd(badint, 1ull); }
    }
}
int main() { unsigned long long *d, h[2] = {0, 0}; cudaMalloc(&d, 16); cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
  k<<<148 * 8, 256>>>(d, d + 1); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("float mismatches %llu, integer-part mismatches %llu\n", h[0], h[1]); return 0; }
