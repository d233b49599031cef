// micro-benchmarks: dependent-chain latencies on one warp (clock64), B200
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_fadd(float *out, float x, long long *cyc) {
    float acc = x; long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) acc = __fadd_rn(acc, x);
    long long t1 = clock64(); out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float *out, float x, long long *cyc) {
    float acc = x; long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) acc = __fmaf_rn(acc, x, x);
    long long t1 = clock64(); out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_fmul_fadd(float *out, const float *in, long long *cyc) {
    float acc = in[0]; float g = in[1], m = in[2]; long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) { float p = __fmul_rn(g, m); acc = __fadd_rn(acc, p); g = __fadd_rn(g, 1.0f); }
    long long t1 = clock64(); out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void k_lds_chain(float *out, const float *in, long long *cyc) {
    __shared__ float4 sm[32 * 64]; __shared__ float4 sg[64];
    for (int i = threadIdx.x; i < 32 * 64; i += 32) sm[i] = make_float4(in[0], in[1], in[2], in[0]);
    for (int i = threadIdx.x; i < 64; i += 32) sg[i] = make_float4(in[1], in[2], in[0], in[1]);
    __syncwarp();
    float acc = 0.f; long long t0 = clock64();
    for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
        for (int g = 0; g < 64; ++g) {
            float4 a = sm[g * 32 + threadIdx.x], b = sg[g];
            acc = __fadd_rn(acc, __fmul_rn(a.x, b.x)); acc = __fadd_rn(acc, __fmul_rn(a.y, b.y));
            acc = __fadd_rn(acc, __fmul_rn(a.z, b.z)); acc = __fadd_rn(acc, __fmul_rn(a.w, b.w));
        }
    }
    long long t1 = clock64(); out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[3] = t1 - t0;
}
int main() {
    float *out, *in; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&in, 64); cudaMallocManaged(&cyc, 64);
    float h[3] = {1.0001f, 0.5f, 1.5f}; cudaMemcpy(in, h, 12, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        k_fadd<<<1, 32>>>(out, 1.0001f, cyc); k_ffma<<<1, 32>>>(out, 1.0001f, cyc);
        k_fmul_fadd<<<1, 32>>>(out, in, cyc); k_lds_chain<<<1, 32>>>(out, in, cyc);
        cudaDeviceSynchronize();
    }
    printf("dependent FADD: %.2f cyc/op\n", cyc[0] / 4096.0);
    printf("dependent FFMA: %.2f cyc/op\n", cyc[1] / 4096.0);
    printf("FMUL+FADD(+FADD) per iter: %.2f cyc\n", cyc[2] / 4096.0);
    printf("LDS.128 x2 + 4x(FMUL,FADD) per row: %.2f cyc/row\n", cyc[3] / (16.0 * 64 * 4));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
