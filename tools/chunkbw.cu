// micro-benchmark: random reads / writes of aligned chunks of 32..1024 bytes over a 16 GiB buffer (each chunk is accessed by
// chunk/16 adjacent lanes with 16-byte vector accesses).  Tells how HBM3e throughput depends on the contiguous run length.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33; return h;
}
template <int LANES, bool WRITE>
__global__ void k(uint4 *buf, size_t n_chunks, unsigned long long *sink, int iters, unsigned long long salt) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t grp = tid / LANES; int lane = tid % LANES;
    unsigned acc = 0;
    unsigned long long h = grp * 0x9E3779B97F4A7C15ull + salt;
    for (int i = 0; i < iters; i++) {
        h = mix(h + i);
        size_t c = h % n_chunks;
        uint4 *p = buf + c * LANES + lane;
        if (WRITE) *p = make_uint4((unsigned)h, lane, i, 3);
        else { uint4 a = *p; acc += a.x + a.y + a.z + a.w; }
    }
    if (!WRITE && acc == 0x12345678u) sink[0] = acc;
}
template <int LANES> void run(uint4 *buf, size_t bytes, unsigned long long *sink) {
    size_t n_chunks = bytes / (16 * LANES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 16, blocks = 148 * 128, threads = 256;
    for (int w = 0; w < 2; w++) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (w) k<LANES, true><<<blocks, threads>>>(buf, n_chunks, sink, iters, rep * 977 + 5);
            else k<LANES, false><<<blocks, threads>>>(buf, n_chunks, sink, iters, rep * 977 + 5);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double bytes_moved = (double)blocks * threads * iters * 16;
        printf("chunk %5d B %s: %.3f ms  %.1f GB/s  %.2f Gchunk/s\n", 16 * LANES, w ? "write" : "read ", best, bytes_moved / best / 1e6,
               bytes_moved / (16 * LANES) / best / 1e6);
    }
}
int main() {
    size_t bytes = (size_t)16 << 30;
    uint4 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    run<2>(buf, bytes, sink); run<4>(buf, bytes, sink); run<8>(buf, bytes, sink); run<16>(buf, bytes, sink); run<32>(buf, bytes, sink);
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
