// ubench_gather.cu - what a B200 delivers for RANDOM 32-byte-sector reads out of a 403 MB buffer
// (3 x the L2): the practical ceiling of the wtosc gather on large sampled waves, where every
// Hermite tap touches its own sector. Each thread issues UNROLL independent 8-byte loads at
// pseudo-random, 2-byte-granular positions per round (like the taps), no arithmetic besides the
// address generator. Reported as sectors x 32 B per second, next to a streaming read of the buffer.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_gather ubench_gather.cu && ./ubench_gather
#include <cstdio>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void gather(const unsigned long long *buf, size_t nwords, int rounds, unsigned long long *sink) {
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned long long acc = 0;
    for (int r = 0; r < rounds; ++r) {
        unsigned long long v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            x = x * 1664525u + 1013904223u;
            v[k] = __ldg(buf + (size_t)(((unsigned long long)x * nwords) >> 32));
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) acc += v[k];
    }
    if (acc == 0x1234567) sink[0] = acc;
}

__global__ void stream(const uint4 *buf, size_t n, unsigned long long *sink) {
    unsigned long long acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(buf + i);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 0x1234567) sink[0] = acc;
}

template <int UNROLL>
static void run(const unsigned long long *buf, size_t nwords, int threads_per_sm, unsigned long long *sink) {
    const int block = 256, grid = 148 * threads_per_sm / block, rounds = 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather<UNROLL><<<grid, block>>>(buf, nwords, 4, sink);
    cudaEventRecord(a);
    gather<UNROLL><<<grid, block>>>(buf, nwords, rounds, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double sectors = (double)grid * block * rounds * UNROLL;
    printf("random 8-byte loads, %4d threads/SM, %2d in flight per thread: %7.1f G sectors/s = %6.0f GB/s of 32-byte sectors\n",
           threads_per_sm, UNROLL, sectors / ms / 1e6, sectors * 32 / ms / 1e6);
}

int main() {
    const size_t bytes = 403ull << 20, nwords = bytes / 8;
    unsigned long long *buf, *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    stream<<<148 * 8, 512>>>((const uint4 *)buf, bytes / 16, sink);
    cudaEventRecord(a);
    stream<<<148 * 8, 512>>>((const uint4 *)buf, bytes / 16, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("streaming read of the same 403 MB: %6.0f GB/s\n", bytes / ms / 1e6);
    run<4>(buf, nwords, 1024, sink);
    run<8>(buf, nwords, 1024, sink);
    run<16>(buf, nwords, 1024, sink);
    run<8>(buf, nwords, 2048, sink);
    run<16>(buf, nwords, 2048, sink);
    run<32>(buf, nwords, 2048, sink);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(err));
    return 0;
}
