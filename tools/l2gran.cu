// micro-benchmark: random 32-byte sector reads over a large buffer under different cudaLimitMaxL2FetchGranularity values
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void rnd_read(const uint4 *buf, size_t n_sectors, unsigned long long *sink, int iters) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long h = tid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int i = 0; i < iters; i++) {
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        size_t s = h % n_sectors;
        uint4 a = buf[2 * s], b = buf[2 * s + 1];
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
__global__ void rnd_write(uint4 *buf, size_t n_sectors, int iters) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long h = tid * 0x9E3779B97F4A7C15ull + 777;
    for (int i = 0; i < iters; i++) {
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        size_t s = h % n_sectors;
        uint4 v = make_uint4((unsigned)h, 1, 2, 3);
        buf[2 * s] = v; buf[2 * s + 1] = v;
    }
}
int main(int argc, char **argv) {
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set %d -> %s\n", gran, cudaGetErrorString(e)); }
    size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity); printf("limit now %zu\n", v);
    size_t bytes = (size_t)8 << 30, n_sectors = bytes / 32;
    uint4 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 16, blocks = 148 * 64, threads = 256;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        rnd_read<<<blocks, threads>>>(buf, n_sectors, sink, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double useful = (double)blocks * threads * iters * 32;
        printf("gran %d read : %.3f ms  useful %.1f GB/s  (%.2f Gsector/s)\n", gran, ms, useful / ms / 1e6, useful / 32 / ms / 1e6);
    }
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        rnd_write<<<blocks, threads>>>(buf, n_sectors, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double useful = (double)blocks * threads * iters * 32;
        printf("gran %d write: %.3f ms  useful %.1f GB/s\n", gran, ms, useful / ms / 1e6);
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
