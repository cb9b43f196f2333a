// dev: dependent-issue latencies on the target (DFMA, rsqrt(double), shared load, barrier)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double* sink, int n) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x * 1e-3 + 1.0;
    __syncthreads();
    double a = sh[threadIdx.x], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    double r = a;
    for (int i = 0; i < n; ++i) r = rsqrt(r + 1.5);
    long long t2 = clock64();
    int idx = threadIdx.x;
    for (int i = 0; i < n; ++i) idx = (int)sh[idx & 1023] & 1023;
    long long t3 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t4 = clock64();
    double q = r;
    for (int i = 0; i < n; ++i) q = 1.0 / (q + 1.5);
    long long t5 = clock64();
    double s = q;
    for (int i = 0; i < n; ++i) s = sqrt(s + 1.5);
    long long t6 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = (t1 - t0) / n; out[1] = (t2 - t1) / n; out[2] = (t3 - t2) / n; out[3] = (t4 - t3) / n;
        out[4] = (t5 - t4) / n; out[5] = (t6 - t5) / n;
    }
    sink[threadIdx.x] = a + r + idx + q + s;
}
int main() {
    long long* o; double* s;
    cudaMalloc(&o, 64); cudaMalloc(&s, 8192);
    for (int threads : {32, 512}) {
        k<<<1, threads>>>(o, s, 1000);
        long long h[6];
        cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %d: dfma %lld rsqrt %lld lds-chain %lld barrier %lld div %lld sqrt %lld cycles\n", threads, h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    return 0;
}
