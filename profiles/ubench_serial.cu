// ubench_serial.cu - cycles per frame of the filter12 recurrence (filter12.c:97-118)
// on ONE warp of an otherwise idle SM, in the formulations render_split could use.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_serial ubench_serial.cu && ./ubench_serial
//
// Lane = voice. The loop reads its input from a shared-memory tile (stride 33, like
// tile A of render_split) and writes its output to another one. Every variant is
// timed with clock64 over kReps x 64 frames; the table prints cycles per frame of the
// slowest participating warp (each participating warp handles 32 x ILP voices).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define DEV __device__ __forceinline__
DEV int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
DEV int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
DEV int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

constexpr int kStride = 33;
constexpr int kFrames = 64;
constexpr int kReps = 64;

struct Coef { int f0, df, q, qstep, lp, bp, hp; };

// V0: the loop as render_split runs it today
template <int ILP>
DEV void loop_base(const int *ta, int *tb, int n, Coef c, int (&d1)[ILP], int (&d2)[ILP]) {
    int f0v = c.f0, qv = c.q;
#pragma unroll 4
    for (int f = 0; f < n; ++f) {
        const int fc = f0v >> 12, qq = qv >> 12;
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const int in = ta[f * kStride + j * kFrames * kStride];
            const int d1s = d1[j] >> 4;
            const int l = wadd(d2[j], wmul(fc, d1s) >> 8);
            const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
            const int bb = wadd(wmul(fc, h >> 4) >> 8, d1[j]);
            tb[f * kStride + j * kFrames * kStride] = wadd(wadd(wmul(l, c.lp), wmul(bb, c.bp)), wmul(h, c.hp)) >> 3;
            d1[j] = bb; d2[j] = l;
        }
        f0v = wadd(f0v, c.df);
        qv = wadd(qv, c.qstep);
    }
}
// V1: coefficients constant over the segment (df == 0, qstep == 0: the common case)
template <int ILP>
DEV void loop_const(const int *ta, int *tb, int n, Coef c, int (&d1)[ILP], int (&d2)[ILP]) {
    const int fc = c.f0 >> 12, qq = c.q >> 12;
#pragma unroll 4
    for (int f = 0; f < n; ++f) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const int in = ta[f * kStride + j * kFrames * kStride];
            const int d1s = d1[j] >> 4;
            const int l = wadd(d2[j], wmul(fc, d1s) >> 8);
            const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
            const int bb = wadd(wmul(fc, h >> 4) >> 8, d1[j]);
            tb[f * kStride + j * kFrames * kStride] = wadd(wadd(wmul(l, c.lp), wmul(bb, c.bp)), wmul(h, c.hp)) >> 3;
            d1[j] = bb; d2[j] = l;
        }
    }
}
// V2: the recurrence warp only publishes (l, b); a helper would finish h and the output mix
template <int ILP>
DEV void loop_pub(const int *ta, int *tb, int n, Coef c, int (&d1)[ILP], int (&d2)[ILP]) {
    int f0v = c.f0, qv = c.q;
#pragma unroll 4
    for (int f = 0; f < n; ++f) {
        const int fc = f0v >> 12, qq = qv >> 12;
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const int in = ta[f * kStride + j * kFrames * kStride];
            const int d1s = d1[j] >> 4;
            const int l = wadd(d2[j], wmul(fc, d1s) >> 8);
            const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
            const int bb = wadd(wmul(fc, h >> 4) >> 8, d1[j]);
            reinterpret_cast<int2 *>(tb)[(f * kStride + j * kFrames * kStride)] = make_int2(l, bb);
            d1[j] = bb; d2[j] = l;
        }
        f0v = wadd(f0v, c.df);
        qv = wadd(qv, c.qstep);
    }
}
// V3: constant coefficients + publish only
template <int ILP>
DEV void loop_const_pub(const int *ta, int *tb, int n, Coef c, int (&d1)[ILP], int (&d2)[ILP]) {
    const int fc = c.f0 >> 12, qq = c.q >> 12;
#pragma unroll 4
    for (int f = 0; f < n; ++f) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const int in = ta[f * kStride + j * kFrames * kStride];
            const int d1s = d1[j] >> 4;
            const int l = wadd(d2[j], wmul(fc, d1s) >> 8);
            const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
            const int bb = wadd(wmul(fc, h >> 4) >> 8, d1[j]);
            reinterpret_cast<int2 *>(tb)[(f * kStride + j * kFrames * kStride)] = make_int2(l, bb);
            d1[j] = bb; d2[j] = l;
        }
    }
}
// V4: pure chain, nothing stored per frame (lower bound of the dependent chain + input load)
template <int ILP>
DEV void loop_chain(const int *ta, int *tb, int n, Coef c, int (&d1)[ILP], int (&d2)[ILP]) {
    const int fc = c.f0 >> 12, qq = c.q >> 12;
#pragma unroll 4
    for (int f = 0; f < n; ++f) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const int in = ta[f * kStride + j * kFrames * kStride];
            const int d1s = d1[j] >> 4;
            const int l = wadd(d2[j], wmul(fc, d1s) >> 8);
            const int h = wsub(wsub(in >> 5, wmul(qq, d1s) >> 8), l);
            const int bb = wadd(wmul(fc, h >> 4) >> 8, d1[j]);
            d1[j] = bb; d2[j] = l;
        }
    }
    (void)tb;
}
// V5: inputs pre-shifted and pre-loaded in registers (8 frames at a time), const coef, publish
template <int ILP>
DEV void loop_regs(const int *ta, int *tb, int n, Coef c, int (&d1)[ILP], int (&d2)[ILP]) {
    const int fc = c.f0 >> 12, qq = c.q >> 12;
    for (int f0 = 0; f0 < n; f0 += 8) {
        int in[ILP][8];
#pragma unroll
        for (int j = 0; j < ILP; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) in[j][k] = ta[(f0 + k) * kStride + j * kFrames * kStride] >> 5;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                const int d1s = d1[j] >> 4;
                const int t = wsub(in[j][k], wmul(qq, d1s) >> 8);
                const int l = wadd(d2[j], wmul(fc, d1s) >> 8);
                const int h = wsub(t, l);
                const int bb = wadd(wmul(fc, h >> 4) >> 8, d1[j]);
                reinterpret_cast<int2 *>(tb)[((f0 + k) * kStride + j * kFrames * kStride)] = make_int2(l, bb);
                d1[j] = bb; d2[j] = l;
            }
        }
    }
}

template <int VAR, int ILP>
__global__ void __launch_bounds__(512) bench(int warp_mask, Coef c, long long *cycles, int *sink, int nfr) {
    extern __shared__ int sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nw = blockDim.x >> 5;
    // one input and one (int2) output tile per warp and ILP slot
    int *ta = sm;     // all warps share one tile set (timing only; races are harmless)
    int *tb = ta + ILP * kFrames * kStride;
    for (int i = lane; i < ILP * kFrames * kStride; i += 32) ta[i] = (i * 2654435761u) >> 8;
    __syncthreads();
    int d1[ILP], d2[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { d1[j] = lane * 977 + j; d2[j] = lane * 131 - j; }
    if (!((warp_mask >> warp) & 1)) return;
    (void)nw;
    c.f0 += lane; c.q += lane * 3;
    long long t0 = clock64();
    for (int r = 0; r < kReps; ++r) {
        if (VAR == 0) loop_base<ILP>(ta + lane, tb + lane, nfr, c, d1, d2);
        if (VAR == 1) loop_const<ILP>(ta + lane, tb + lane, nfr, c, d1, d2);
        if (VAR == 2) loop_pub<ILP>(ta + lane, tb + 2 * lane, nfr, c, d1, d2);
        if (VAR == 3) loop_const_pub<ILP>(ta + lane, tb + 2 * lane, nfr, c, d1, d2);
        if (VAR == 4) loop_chain<ILP>(ta + lane, tb + lane, nfr, c, d1, d2);
        if (VAR == 5) loop_regs<ILP>(ta + lane, tb + 2 * lane, nfr, c, d1, d2);
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
    int s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += d1[j] + d2[j];
    sink[blockIdx.x * blockDim.x + tid] = s + tb[lane];
}

template <int VAR, int ILP>
static void run(const char *name, int warps, int warp_mask, Coef c) {
    long long *d_cyc;
    int *d_sink;
    const int grid = 148;
    cudaMalloc(&d_cyc, grid * 16 * sizeof(long long));
    cudaMalloc(&d_sink, grid * 512 * sizeof(int));
    cudaMemset(d_cyc, 0, grid * 16 * sizeof(long long));
    size_t smem = (size_t)ILP * kFrames * kStride * 3 * sizeof(int);
    cudaFuncSetAttribute(bench<VAR, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 3; ++it) bench<VAR, ILP><<<grid, warps * 32, smem>>>(warp_mask, c, d_cyc, d_sink, kFrames);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%-44s CUDA error %s\n", name, cudaGetErrorString(err)); return; }
    long long h[148 * 16];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double worst = 0, sum = 0; int cnt = 0;
    for (int b = 0; b < grid; ++b)
        for (int w = 0; w < 16; ++w)
            if (h[b * 16 + w]) { double x = (double)h[b * 16 + w] / (kReps * kFrames); worst = x > worst ? x : worst; sum += x; ++cnt; }
    int nact = __builtin_popcount(warp_mask);
    printf("%-44s ILP %d warps %d: %6.1f cyc/frame (avg %6.1f) -> %5.2f cyc per frame and 32 voices\n", name, ILP, nact,
           worst, sum / cnt, worst / (ILP * nact));
    cudaFree(d_cyc); cudaFree(d_sink);
}

// Contention: the recurrence warp (warp 3, alone on sub-partition 3) runs V0 while `nload` warps on
// the other sub-partitions do what render_split's helpers do in stage A - data-dependent 16-byte
// gathers from an 80 KB table in shared memory - back to back.
__global__ void __launch_bounds__(512) bench_contended(int nload, Coef c, long long *cycles, int *sink, int nfr) {
    extern __shared__ int sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int *ta = sm, *tb = sm + kFrames * kStride;
    int4 *table = reinterpret_cast<int4 *>(sm + 3 * kFrames * kStride);
    const int tabn = 5000;
    for (int i = tid; i < tabn; i += blockDim.x) table[i] = make_int4(i, i * 3, i * 5, i * 7);
    for (int i = tid; i < kFrames * kStride; i += blockDim.x) ta[i] = (i * 2654435761u) >> 8;
    __shared__ volatile int stop;
    if (tid == 0) stop = 0;
    __syncthreads();
    if (warp == 3) {
        int d1[1] = {lane * 977}, d2[1] = {lane * 131};
        c.f0 += lane; c.q += lane * 3;
        long long t0 = clock64();
        for (int r = 0; r < kReps; ++r) loop_base<1>(ta + lane, tb + lane, nfr, c, d1, d2);
        long long t1 = clock64();
        if (lane == 0) { cycles[blockIdx.x * 16 + warp] = t1 - t0; stop = 1; }
        sink[blockIdx.x * blockDim.x + tid] = d1[0] + d2[0];
        return;
    }
    const int q = (warp & 3) == 3 ? -1 : warp - (warp >> 2);
    if (q < 0 || q >= nload) return;
    unsigned x = tid * 2654435761u + 1;
    int acc = 0;
    while (!stop) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {          // 6 frames x 2 taps, like one helper's slice
            x = x * 1664525u + 1013904223u;
            const int4 e = table[(x >> 8) % tabn];
            acc += e.x + ((e.y * (int)(x & 0x7fff)) >> 15) + e.z + e.w;
        }
    }
    sink[blockIdx.x * blockDim.x + tid] = acc;
}

static void run_contended(int nload, Coef c) {
    long long *d_cyc;
    int *d_sink;
    const int grid = 148;
    cudaMalloc(&d_cyc, grid * 16 * sizeof(long long));
    cudaMalloc(&d_sink, grid * 512 * sizeof(int));
    cudaMemset(d_cyc, 0, grid * 16 * sizeof(long long));
    size_t smem = (size_t)3 * kFrames * kStride * sizeof(int) + 5000 * 16;
    cudaFuncSetAttribute(bench_contended, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 3; ++it) bench_contended<<<grid, 512, smem>>>(nload, c, d_cyc, d_sink, kFrames);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("contended: CUDA error %s\n", cudaGetErrorString(err)); return; }
    long long h[148 * 16];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double worst = 0, sum = 0; int cnt = 0;
    for (int b = 0; b < grid; ++b)
        if (h[b * 16 + 3]) { double v = (double)h[b * 16 + 3] / (kReps * kFrames); worst = v > worst ? v : worst; sum += v; ++cnt; }
    printf("V0 on warp 3 with %2d warps gathering from shared memory on sub-partitions 0-2: %6.1f cyc/frame (avg %6.1f)\n",
           nload, worst, sum / cnt);
    cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
    Coef ramp = {3000 << 12, 37, 9000 << 12, 11, 256, 3, 5};
    Coef flat = {3000 << 12, 0, 9000 << 12, 0, 256, 0, 0};
    const int one = 1 << 3;             // warp 3 alone on sub-partition 3
    run<0, 1>("V0 today: ramping coef, mix on the warp", 4, one, ramp);
    run<1, 1>("V1 constant coef, mix on the warp", 4, one, flat);
    run<2, 1>("V2 ramping coef, publish (l,b) only", 4, one, ramp);
    run<3, 1>("V3 constant coef, publish (l,b) only", 4, one, flat);
    run<4, 1>("V4 pure chain (no store)", 4, one, flat);
    run<5, 1>("V5 const, inputs in regs, publish", 4, one, flat);
    run<0, 2>("V0 x2 voices per lane", 4, one, ramp);
    run<1, 2>("V1 x2 voices per lane", 4, one, flat);
    run<3, 2>("V3 x2 voices per lane", 4, one, flat);
    run<5, 2>("V5 x2 voices per lane", 4, one, flat);
    run<3, 3>("V3 x3 voices per lane", 4, one, flat);
    run<3, 4>("V3 x4 voices per lane", 4, one, flat);
    run<5, 4>("V5 x4 voices per lane", 4, one, flat);
    // two / four warps: same sub-partition (3, 7, 11, 15) or spread (0..3)
    run<0, 1>("V0 two warps, same sub-partition", 8, (1 << 3) | (1 << 7), ramp);
    run<0, 1>("V0 two warps, different sub-partitions", 4, (1 << 3) | (1 << 2), ramp);
    run<3, 1>("V3 two warps, same sub-partition", 8, (1 << 3) | (1 << 7), flat);
    run<3, 1>("V3 four warps, same sub-partition", 16, (1 << 3) | (1 << 7) | (1 << 11) | (1 << 15), flat);
    run<3, 1>("V3 four warps, one per sub-partition", 4, 0xf, flat);
    run<3, 2>("V3 x2, four warps, one per sub-partition", 4, 0xf, flat);
    run<3, 1>("V3 eight warps, two per sub-partition", 8, 0xff, flat);
    run<3, 1>("V3 sixteen warps, four per sub-partition", 16, 0xffff, flat);
    run<0, 1>("V0 sixteen warps, four per sub-partition", 16, 0xffff, ramp);
    for (int n : {0, 3, 6, 11}) run_contended(n, ramp);
    return 0;
}
